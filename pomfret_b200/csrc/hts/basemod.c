/* hts-shim base-modification tags (SAMtags "Base modifications": MM/Mm, ML/Ml,
 * MN).  API shape of htslib's bam_parse_basemod / bam_mods_at_next_pos as
 * used at reference blockjoin.c:807,833.
 *
 * Implementation is eager: the MM string is resolved to a position-sorted
 * event list once, then bam_mods_at_next_pos() hands out the events of the
 * next SEQ position.  Semantics restated from the SAMtags specification
 * (SURVEY.md App. A.1):
 *   - a segment is  BASE STRAND CODES [.?] (,delta)* ;
 *   - deltas count occurrences of BASE on the *original* read strand to skip;
 *     for reverse-strand alignments SEQ is reverse complemented, so the walk
 *     runs over the complement base from the other end;
 *   - several codes in one segment share positions, ML interleaved;
 *   - segments consume ML in tag order; when ML is present its length must
 *     equal the number of (delta, code) pairs, otherwise the record carries
 *     no modifications;
 *   - a reverse-strand list that asks for more bases than SEQ holds is an
 *     error (no modifications); on the forward strand the excess is silently
 *     never reached.
 * Any malformed tag => no modifications for the record (parse returns -1).
 */
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "htslib/sam.h"

#define MAX_BASE_MOD 256

typedef struct {
    int32_t pos;
    uint32_t order;
    hts_base_mod m;
} mod_event_t;

struct hts_base_mod_state {
    int n_events, m_events;
    mod_event_t *events;
    int cursor;   /* next event to hand out */
    int seq_pos;  /* next SEQ position for bam_mods_at_next_pos */
    int nmods;    /* number of (segment, code) streams */
    int is_rev, l_qseq;
};

hts_base_mod_state *hts_base_mod_state_alloc(void) {
    return (hts_base_mod_state *)calloc(1, sizeof(hts_base_mod_state));
}
void hts_base_mod_state_free(hts_base_mod_state *s) {
    if (!s) return;
    free(s->events);
    free(s);
}

static const uint8_t k_rc16[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};

static void push_event(hts_base_mod_state *s, int pos, const hts_base_mod *m) {
    if (s->n_events == s->m_events) {
        s->m_events = s->m_events ? s->m_events * 2 : 256;
        s->events = (mod_event_t *)realloc(s->events, sizeof(mod_event_t) * s->m_events);
    }
    s->events[s->n_events].pos = pos;
    s->events[s->n_events].order = (uint32_t)s->n_events;
    s->events[s->n_events].m = *m;
    s->n_events++;
}

static int cmp_event(const void *a, const void *b) {
    const mod_event_t *x = (const mod_event_t *)a, *y = (const mod_event_t *)b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return x->order < y->order ? -1 : x->order > y->order ? 1 : 0;
}

static int parse_fail(hts_base_mod_state *s) {
    s->n_events = 0;
    s->nmods = 0;
    return -1;
}

int bam_parse_basemod(const bam1_t *b, hts_base_mod_state *s) {
    s->n_events = 0;
    s->cursor = 0;
    s->seq_pos = 0;
    s->nmods = 0;
    s->is_rev = (b->core.flag & BAM_FREVERSE) != 0;
    s->l_qseq = b->core.l_qseq;

    uint8_t *mm = bam_aux_get(b, "MM");
    if (!mm) mm = bam_aux_get(b, "Mm");
    if (!mm) return 0;
    if (mm[0] != 'Z') return -1;

    uint8_t *mn = bam_aux_get(b, "MN");
    if (mn && bam_aux2i(mn) != b->core.l_qseq && b->core.l_qseq) return -1;

    uint8_t *ml = bam_aux_get(b, "ML");
    if (!ml) ml = bam_aux_get(b, "Ml");
    if (ml && (ml[0] != 'B' || ml[1] != 'C')) return -1;
    int64_t ml_len = 0;
    if (ml) {
        ml_len = (int64_t)ml[2] | ((int64_t)ml[3] << 8) | ((int64_t)ml[4] << 16) | ((int64_t)ml[5] << 24);
        ml += 6;
    }

    const uint8_t *seq = bam_get_seq(b);
    const int len = b->core.l_qseq;
    const int rev = s->is_rev;

    int freq[16] = {0};
    if (rev)
        for (int i = 0; i < len; i++) freq[bam_seqi(seq, i)]++;

    const char *cp = (const char *)mm + 1;
    int64_t ml_used = 0;
    int mod_num = 0;
    int64_t *cum = NULL; /* scratch: cumulative (delta+1) */
    size_t cum_m = 0;

    while (*cp) {
        /* header */
        unsigned char bchar = (unsigned char)*cp++;
        if (bchar != 'A' && bchar != 'C' && bchar != 'G' && bchar != 'T' && bchar != 'U' && bchar != 'N') goto fail;
        if (bchar == 'U') bchar = 'T';
        int bcode = seq_nt16_table[bchar];
        if (*cp != '+' && *cp != '-') goto fail;
        int strand = (*cp++ == '-');
        int codes[MAX_BASE_MOD];
        int n_codes = 0;
        if (isdigit((unsigned char)*cp)) {
            char *e;
            long chebi = strtol(cp, &e, 10);
            cp = e;
            codes[n_codes++] = -(int)chebi;
        } else {
            while (*cp && isalpha((unsigned char)*cp)) {
                if (n_codes < MAX_BASE_MOD) codes[n_codes] = (unsigned char)*cp;
                n_codes++;
                cp++;
            }
            if (*cp == '\0') goto fail;
        }
        if (*cp == '.' || *cp == '?') cp++;
        else if (*cp != ',' && *cp != ';') goto fail;
        if (mod_num + n_codes >= MAX_BASE_MOD && n_codes > 0) goto fail;
        const int stride = n_codes;

        /* delta list */
        size_t n_delta = 0;
        int64_t total = 0;
        while (*cp == ',') {
            char *e;
            long d = strtol(cp + 1, &e, 10);
            if (e == cp + 1) goto fail;
            cp = e;
            if (n_delta == cum_m) {
                cum_m = cum_m ? cum_m * 2 : 512;
                cum = (int64_t *)realloc(cum, sizeof(int64_t) * cum_m);
            }
            total += (int64_t)d + 1;
            cum[n_delta++] = total;
        }
        if (*cp != ';') goto fail; /* missing semicolon or junk */
        cp++;

        if (ml) {
            if (ml_used + (int64_t)n_delta * stride > ml_len) goto fail;
        }

        /* resolve positions: index among matching bases, counted from the left of SEQ */
        int64_t lead = 0; /* reverse: matching bases skipped at the left edge */
        if (rev) {
            lead = (int64_t)freq[k_rc16[bcode]] - total;
            if (lead < 0 && n_codes > 0) goto fail;
        }
        if (n_delta > 0 && n_codes > 0) {
            /* walk SEQ once, handing out targets in increasing left-index order */
            size_t k = rev ? n_delta : 0; /* rev: next target is k-1 */
            int64_t match_idx = 0;
            for (int p = 0; p < len; p++) {
                int code = bam_seqi(seq, p);
                if (rev) code = k_rc16[code];
                if (code != bcode && bcode != 15) continue;
                int64_t want;
                size_t which;
                if (!rev) {
                    if (k >= n_delta) break;
                    which = k;
                    want = cum[k] - 1;
                } else {
                    if (k == 0) break;
                    which = k - 1;
                    want = lead + (total - cum[which]);
                }
                if (match_idx == want) {
                    for (int c = 0; c < n_codes; c++) {
                        hts_base_mod m;
                        m.modified_base = codes[c];
                        m.canonical_base = seq_nt16_str[bcode];
                        m.strand = strand;
                        m.qual = ml ? ml[ml_used + (int64_t)which * stride + c] : HTS_MOD_UNKNOWN;
                        push_event(s, p, &m);
                    }
                    if (!rev) k++; else k--;
                }
                match_idx++;
            }
        }
        ml_used += (int64_t)n_delta * stride;
        mod_num += n_codes;
    }
    if (ml && ml_used != ml_len) goto fail;
    free(cum);
    s->nmods = mod_num;
    /* events of one segment are position sorted; interleave segments */
    qsort(s->events, s->n_events, sizeof(mod_event_t), cmp_event);
    return 0;
fail:
    free(cum);
    return parse_fail(s);
}

int bam_mods_at_next_pos(const bam1_t *b, hts_base_mod_state *s, hts_base_mod *mods, int n_mods) {
    if (!s->is_rev && s->seq_pos >= b->core.l_qseq) return -1;
    int pos = s->seq_pos++;
    int n = 0;
    while (s->cursor < s->n_events && s->events[s->cursor].pos < pos) s->cursor++;
    while (s->cursor < s->n_events && s->events[s->cursor].pos == pos) {
        if (n < n_mods) mods[n] = s->events[s->cursor].m;
        n++;
        s->cursor++;
    }
    return n;
}

int bam_next_basemod(const bam1_t *b, hts_base_mod_state *s, hts_base_mod *mods, int n_mods, int *pos) {
    (void)b;
    if (s->cursor >= s->n_events) return 0;
    int p = s->events[s->cursor].pos;
    int n = 0;
    while (s->cursor < s->n_events && s->events[s->cursor].pos == p) {
        if (n < n_mods) mods[n] = s->events[s->cursor].m;
        n++;
        s->cursor++;
    }
    s->seq_pos = p + 1;
    *pos = p;
    return n;
}
