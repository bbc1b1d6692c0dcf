/* TEST INFRASTRUCTURE (oracle port) — per-read decode.
 *
 * Restates, for one packed alignment record:
 *   - htslib's bam_parse_basemod / bam_mods_at_next_pos as the reference uses them
 *     (blockjoin.c:807, 832-882).  htslib is not in /root/reference; semantics restated from the
 *     SAMtags specification, see SURVEY.md App. A.1 ("parity unpinned" for this layer);
 *   - the 5mC / CpG filter and ML categorisation of fill_read_meth_record_from_bam_line
 *     (blockjoin.c:794-908);
 *   - get_mod_poss_on_ref (blockjoin.c:605-792) including its quirks;
 *   - bam_endpos.
 */
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "port.h"

static inline int nib(const uint8_t *seq, uint32_t i) { return (seq[i >> 1] >> ((~i & 1) << 2)) & 0xf; }

static void calls_push(port_calls_t *c, uint32_t pos, uint8_t cat) {
    if (c->n == c->m) {
        c->m = c->m ? c->m * 2 : 64;
        c->pos = (uint32_t *)realloc(c->pos, sizeof(uint32_t) * c->m);
        c->cat = (uint8_t *)realloc(c->cat, c->m);
    }
    c->pos[c->n] = pos;
    c->cat[c->n] = cat;
    c->n++;
}

/* ---------------- MM / ML ---------------- */

typedef struct {
    int32_t pos;
    uint32_t ord;
    int code;   /* >0: ASCII code letter, <=0: -ChEBI */
    int canon;  /* 4-bit base code of the canonical base as written in MM */
    int qual;   /* ML byte or -1 */
} mod_ev_t;

typedef struct { mod_ev_t *a; size_t n, m; } ev_vec_t;

static void ev_push(ev_vec_t *v, int32_t pos, int code, int canon, int qual) {
    if (v->n == v->m) {
        v->m = v->m ? v->m * 2 : 256;
        v->a = (mod_ev_t *)realloc(v->a, sizeof(mod_ev_t) * v->m);
    }
    v->a[v->n].pos = pos; v->a[v->n].ord = (uint32_t)v->n;
    v->a[v->n].code = code; v->a[v->n].canon = canon; v->a[v->n].qual = qual;
    v->n++;
}
static int ev_cmp(const void *a, const void *b) {
    const mod_ev_t *x = (const mod_ev_t *)a, *y = (const mod_ev_t *)b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return x->ord < y->ord ? -1 : x->ord > y->ord;
}

static int base_code(int ch) {
    switch (ch) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': case 'U': return 8; case 'N': return 15;
    default: return -1;
    }
}
static const uint8_t k_comp[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};

/* Resolve the MM/ML tags to (SEQ position, code, canonical, qual) events, position sorted, ties in
 * tag order.  Returns 0, or -1 when the tags are malformed (=> the record has no modifications). */
static int resolve_basemods(const pomfret_gpu_read_desc *r, ev_vec_t *ev) {
    ev->n = 0;
    if (!r->mm) return 0;
    if (r->tags_malformed) return -1;
    if (r->mn >= 0 && (uint32_t)r->mn != r->l_qseq && r->l_qseq) return -1;
    const int rev = (r->flag & 16) != 0;
    const int len = (int)r->l_qseq;
    int freq[16] = {0};
    if (rev) for (int i = 0; i < len; i++) freq[nib(r->seq, (uint32_t)i)]++;
    const char *s = r->mm, *e = r->mm + r->mm_len;
    /* htslib parses a NUL terminated string: stop at an embedded NUL */
    for (const char *q = s; q < e; q++) if (!*q) { e = q; break; }
    int64_t ml_used = 0;
    int n_streams = 0;
    int64_t *cum = NULL;
    size_t cum_m = 0;
    while (s < e) {
        int canon = base_code((unsigned char)*s++);
        if (canon < 0) goto bad;
        if (s >= e || (*s != '+' && *s != '-')) goto bad;
        s++;
        int codes[256], n_codes = 0;
        if (s < e && isdigit((unsigned char)*s)) {
            long v = 0;
            while (s < e && isdigit((unsigned char)*s)) v = v * 10 + (*s++ - '0');
            codes[n_codes++] = -(int)v;
        } else {
            while (s < e && isalpha((unsigned char)*s)) {
                if (n_codes < 256) codes[n_codes] = (unsigned char)*s;
                n_codes++; s++;
            }
            if (s >= e) goto bad;
        }
        if (s < e && (*s == '.' || *s == '?')) s++;
        else if (s >= e || (*s != ',' && *s != ';')) goto bad;
        if (n_codes > 0 && n_streams + n_codes >= 256) goto bad;
        size_t nd = 0;
        int64_t total = 0;
        while (s < e && *s == ',') {
            s++;
            if (s >= e || !isdigit((unsigned char)*s)) goto bad;
            int64_t v = 0;
            while (s < e && isdigit((unsigned char)*s)) v = v * 10 + (*s++ - '0');
            if (nd == cum_m) { cum_m = cum_m ? cum_m * 2 : 512; cum = (int64_t *)realloc(cum, sizeof(int64_t) * cum_m); }
            total += v + 1;
            cum[nd++] = total;
        }
        if (s >= e || *s != ';') goto bad;
        s++;
        if (r->ml_len >= 0 && ml_used + (int64_t)nd * n_codes > r->ml_len) goto bad;
        int64_t lead = 0;
        if (rev) {
            lead = (int64_t)freq[k_comp[canon]] - total;
            if (lead < 0 && n_codes > 0) goto bad;
        }
        if (nd > 0 && n_codes > 0) {
            /* k-th listed base sits on the (cum[k]-1)-th matching base counted from the read's 5' end;
             * for reverse alignments SEQ is reverse-complemented, so count from the right end of SEQ,
             * which equals (lead + total - cum[k]) from the left. */
            int64_t midx = 0;
            size_t k = rev ? nd : 0;
            for (int p = 0; p < len; p++) {
                int c = nib(r->seq, (uint32_t)p);
                if (rev) c = k_comp[c];
                if (c != canon && canon != 15) continue;
                size_t which;
                int64_t want;
                if (!rev) { if (k >= nd) break; which = k; want = cum[k] - 1; }
                else { if (k == 0) break; which = k - 1; want = lead + (total - cum[which]); }
                if (midx == want) {
                    for (int c2 = 0; c2 < n_codes; c2++)
                        ev_push(ev, p, codes[c2], canon,
                                r->ml_len >= 0 ? r->ml[ml_used + (int64_t)which * n_codes + c2] : -1);
                    if (!rev) k++; else k--;
                }
                midx++;
            }
        }
        ml_used += (int64_t)nd * n_codes;
        n_streams += n_codes;
    }
    if (r->ml_len >= 0 && ml_used != r->ml_len) goto bad;
    free(cum);
    qsort(ev->a, ev->n, sizeof(mod_ev_t), ev_cmp);
    return 0;
bad:
    free(cum);
    ev->n = 0;
    return -1;
}

/* ---------------- get_mod_poss_on_ref ---------------- */

static void push_or_overwrite(port_calls_t *out, uint32_t pos, uint8_t cat) {
    /* blockjoin.c:704-709 */
    if (out->n > 0 && out->pos[out->n - 1] == pos) out->cat[out->n - 1] = cat;
    else calls_push(out, pos, cat);
}

static void implicit_fill(port_calls_t *out, const uint8_t *seq, uint32_t l_qseq, uint32_t from, uint32_t until,
                          uint32_t i_ref, int offset) {
    /* blockjoin.c:670-699 / 731-760: every CG in SEQ[from, until) becomes an unmethylated call */
    for (uint32_t t = from; t < until; t++) {
        if (t < l_qseq - 1 && nib(seq, t) == 2 && nib(seq, t + 1) == 4) {
            uint32_t p = i_ref + t + (uint32_t)offset;
            if (!(out->n > 0 && out->pos[out->n - 1] == p)) calls_push(out, p, 1);
            t++;
        }
    }
}

int port_map_mods_to_ref(const uint32_t *cigar, int n_cigar, uint32_t qs, int strand, const uint32_t *mod_pos,
                         const uint8_t *mod_cat, int n_mods, const uint8_t *seq, uint32_t l_qseq,
                         port_calls_t *out) {
    if (n_cigar == 0 || n_mods == 0) return 0; /* :614 */
    const int cg = strand ? -1 : 0;            /* :617-618 */
    uint32_t i_read = 0, i_ref = qs;
    uint32_t it = 0, next = mod_pos[0];
    uint8_t nq = mod_cat[0];
    int ic = 0;
    if ((cigar[0] & 15) == 4) { /* leading soft clip, :629-652 */
        i_read = cigar[0] >> 4;
        while (next < i_read) {
            it++;
            if (it < (uint32_t)n_mods) { next = mod_pos[it]; nq = mod_cat[it]; }
            else break;
        }
        if (next == i_read) {
            calls_push(out, i_ref + (uint32_t)cg, nq);
            it++;
            if (it < (uint32_t)n_mods) { next = mod_pos[it]; nq = mod_cat[it]; }
        }
        i_ref -= cigar[0] >> 4;
        ic = 1;
    }
    int offset = 0;
    for (; ic < n_cigar; ic++) {
        uint32_t op = cigar[ic] & 15, L = cigar[ic] >> 4;
        if (op <= 1) {
            uint32_t pos_canonical = i_read;
            while (i_read + L >= next) { /* inclusive, :663 */
                if (op == 0 && next != UINT32_MAX) {
                    if (seq) {
                        uint32_t until = next - 1 < i_read + L ? next - 1 : i_read + L;
                        implicit_fill(out, seq, l_qseq, pos_canonical, until, i_ref, offset);
                    }
                    push_or_overwrite(out, i_ref + next + (uint32_t)cg + (uint32_t)offset, nq);
                    pos_canonical = cg == 0 ? next + 1 : next + 2;
                }
                it++;
                if (it >= (uint32_t)n_mods) { next = UINT32_MAX; break; }
                next = mod_pos[it];
                nq = mod_cat[it];
            }
            if (op == 0) {
                if (seq) implicit_fill(out, seq, l_qseq, pos_canonical, i_read + L, i_ref, offset);
                i_read += L;
            } else {
                i_read += L;
                offset -= (int)L;
            }
        } else if (op == 2) offset += (int)L;
        else if (op == 3) break;
        else if (op == 4) break;
        else return POMFRET_GPU_ERR_FATAL_CIGAR; /* :776-778 exit(1) */
    }
    return 1;
}

/* ---------------- whole read ---------------- */

int port_decode_read(const pomfret_gpu_read_desc *r, int lo, int hi, port_calls_t *out, uint32_t *status,
                     uint32_t *end_pos) {
    out->n = 0;
    uint32_t st = 0;
    /* bam_endpos */
    uint64_t rlen = 0;
    for (uint32_t i = 0; i < r->n_cigar; i++) {
        uint32_t op = r->cigar[i] & 15;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += r->cigar[i] >> 4;
    }
    if (r->flag & 4) rlen = 0;
    if (rlen == 0) rlen = 1;
    if (end_pos) *end_pos = (uint32_t)(r->pos + rlen);

    ev_vec_t ev = {0, 0, 0};
    if (resolve_basemods(r, &ev) != 0) st |= POMFRET_GPU_READ_MM_ERROR;

    /* blockjoin.c:832-882: walk SEQ positions, keep C+m at CpG */
    const uint32_t len = r->l_qseq;
    const uint8_t qlo = (uint8_t)lo, qhi = (uint8_t)hi;
    uint32_t *mp = NULL;
    uint8_t *mc = NULL;
    size_t nm = 0, mm_cap = 0;
    int has_implicit = 0;
    for (size_t i = 0; i < ev.n;) {
        size_t j = i;
        while (j < ev.n && ev.a[j].pos == ev.a[i].pos) j++;
        size_t n_here = j - i;
        uint32_t p = (uint32_t)ev.a[i].pos;
        if (n_here <= PORT_N_MODS) { /* n > N_MODS: warning + position skipped, :838-843 */
            for (size_t t = i; t < j; t++) {
                if (ev.a[t].canon == 2 && ev.a[t].code == 'm' && p < len - 1 && p > 0) {
                    int ok = nib(r->seq, p) == 2 ? nib(r->seq, p + 1) == 4 : nib(r->seq, p - 1) == 2;
                    if (!ok) { has_implicit = 1; continue; }
                    if (nm == mm_cap) {
                        mm_cap = mm_cap ? mm_cap * 2 : 256;
                        mp = (uint32_t *)realloc(mp, sizeof(uint32_t) * (mm_cap + 1));
                        mc = (uint8_t *)realloc(mc, mm_cap + 1);
                    }
                    uint8_t q = (uint8_t)ev.a[t].qual;
                    mp[nm] = p;
                    mc[nm] = q < qlo ? 1 : q >= qhi ? 0 : 2;
                    nm++;
                }
            }
        }
        i = j;
    }
    free(ev.a);
    if (has_implicit) st |= POMFRET_GPU_READ_HAS_IMPLICIT;
    int rc = port_map_mods_to_ref(r->cigar, (int)r->n_cigar, r->pos, (r->flag & 16) != 0, mp, mc, (int)nm,
                                  has_implicit ? r->seq : NULL, len, out);
    free(mp);
    free(mc);
    if (rc == POMFRET_GPU_ERR_FATAL_CIGAR) st |= POMFRET_GPU_READ_FATAL_CIGAR;
    else if (rc == 1) st |= POMFRET_GPU_READ_KEPT;
    if (status) *status = st;
    return rc;
}
