/* hts-shim BAM records, header, aux fields (SAMv1 §4.2). */
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <stdio.h>
#include "htslib/sam.h"

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
const unsigned char seq_nt16_table[256] = {
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
     1,  2,  4,  8, 15, 15, 15, 15, 15, 15, 15, 15, 15,  0, 15, 15,
    15,  1, 14,  2, 13, 15, 15,  4, 11, 15, 15, 12, 15,  3, 15, 15,
    15, 15,  5,  6,  8, 15,  7,  9, 15, 10, 15, 15, 15, 15, 15, 15,
    15,  1, 14,  2, 13, 15, 15,  4, 11, 15, 15, 12, 15,  3, 15, 15,
    15, 15,  5,  6,  8, 15,  7,  9, 15, 10, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
    15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15,
};

static uint64_t g_n_records, g_n_bytes;
void hts_shim_counters(uint64_t *n_records, uint64_t *n_bytes) {
    if (n_records) *n_records = g_n_records;
    if (n_bytes) *n_bytes = g_n_bytes;
}

static inline uint32_t ld32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint16_t ld16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline void st32(uint8_t *p, uint32_t v) {
    p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = (v >> 24) & 0xff;
}
static inline void st16(uint8_t *p, uint16_t v) { p[0] = v & 0xff; p[1] = v >> 8; }

/* ---------- files ---------- */

htsFile *hts_open(const char *fn, const char *mode) {
    int is_write = strchr(mode, 'w') != NULL;
    BGZF *bg = bgzf_open(fn, mode);
    if (!bg) return NULL;
    htsFile *fp = (htsFile *)calloc(1, sizeof(htsFile));
    fp->is_bin = 1;
    fp->is_write = is_write;
    fp->is_bgzf = 1;
    fp->fn = strdup(fn);
    fp->fp.bgzf = bg;
    if (!is_write) {
        /* sniff: a BGZF member must start the file */
        uint8_t magic[2];
        FILE *f = fopen(fn, "rb");
        int ok = f && fread(magic, 1, 2, f) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        if (f) fclose(f);
        if (!ok) fp->is_bgzf = 0;
    }
    return fp;
}

int hts_close(htsFile *fp) {
    if (!fp) return -1;
    int rc = bgzf_close(fp->fp.bgzf);
    free(fp->fn);
    free(fp);
    return rc;
}

/* ---------- header ---------- */

sam_hdr_t *bam_hdr_read(BGZF *fp) {
    uint8_t buf[8];
    if (bgzf_read(fp, buf, 4) != 4 || memcmp(buf, "BAM\1", 4) != 0) return NULL;
    if (bgzf_read(fp, buf, 4) != 4) return NULL;
    sam_hdr_t *h = (sam_hdr_t *)calloc(1, sizeof(sam_hdr_t));
    h->l_text = ld32(buf);
    h->text = (char *)malloc(h->l_text + 1);
    if (bgzf_read(fp, h->text, h->l_text) != (ssize_t)h->l_text) goto fail;
    h->text[h->l_text] = 0;
    if (bgzf_read(fp, buf, 4) != 4) goto fail;
    h->n_targets = (int32_t)ld32(buf);
    h->target_name = (char **)calloc(h->n_targets > 0 ? h->n_targets : 1, sizeof(char *));
    h->target_len = (uint32_t *)calloc(h->n_targets > 0 ? h->n_targets : 1, sizeof(uint32_t));
    for (int i = 0; i < h->n_targets; i++) {
        if (bgzf_read(fp, buf, 4) != 4) goto fail;
        uint32_t l = ld32(buf);
        h->target_name[i] = (char *)malloc(l + 1);
        if (bgzf_read(fp, h->target_name[i], l) != (ssize_t)l) goto fail;
        h->target_name[i][l] = 0;
        if (bgzf_read(fp, buf, 4) != 4) goto fail;
        h->target_len[i] = ld32(buf);
    }
    return h;
fail:
    sam_hdr_destroy(h);
    return NULL;
}

sam_hdr_t *sam_hdr_read(samFile *fp) {
    if (!fp || !fp->is_bgzf) return NULL;
    if (bgzf_seek(fp->fp.bgzf, 0, SEEK_SET) != 0) return NULL;
    return bam_hdr_read(fp->fp.bgzf);
}

int bam_hdr_write(BGZF *fp, const sam_hdr_t *h) {
    uint8_t buf[8];
    if (bgzf_write(fp, "BAM\1", 4) != 4) return -1;
    st32(buf, (uint32_t)h->l_text);
    if (bgzf_write(fp, buf, 4) != 4) return -1;
    if (h->l_text && bgzf_write(fp, h->text, h->l_text) != (ssize_t)h->l_text) return -1;
    st32(buf, (uint32_t)h->n_targets);
    if (bgzf_write(fp, buf, 4) != 4) return -1;
    for (int i = 0; i < h->n_targets; i++) {
        uint32_t l = (uint32_t)strlen(h->target_name[i]) + 1;
        st32(buf, l);
        if (bgzf_write(fp, buf, 4) != 4) return -1;
        if (bgzf_write(fp, h->target_name[i], l) != (ssize_t)l) return -1;
        st32(buf, h->target_len[i]);
        if (bgzf_write(fp, buf, 4) != 4) return -1;
    }
    if (bgzf_flush(fp) != 0) return -1;
    return 0;
}

void sam_hdr_destroy(sam_hdr_t *h) {
    if (!h) return;
    if (h->target_name) {
        for (int i = 0; i < h->n_targets; i++) free(h->target_name[i]);
        free(h->target_name);
    }
    free(h->target_len);
    free(h->text);
    free(h);
}

int sam_hdr_name2tid(sam_hdr_t *h, const char *ref) {
    for (int i = 0; i < h->n_targets; i++)
        if (strcmp(h->target_name[i], ref) == 0) return i;
    return -1;
}

/* ---------- records ---------- */

bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }

void bam_destroy1(bam1_t *b) {
    if (!b) return;
    free(b->data);
    free(b);
}

static int reserve(bam1_t *b, size_t want) {
    if (want <= b->m_data) return 0;
    if (b->id == POMFRET_BAM_EXTERNAL_DATA) return -1; /* caller-owned buffer: never reallocated */
    size_t m = b->m_data ? b->m_data : 256;
    while (m < want) m += m >> 1;
    uint8_t *d = (uint8_t *)realloc(b->data, m);
    if (!d) return -1;
    b->data = d;
    b->m_data = (uint32_t)m;
    return 0;
}

/* Alignments with more than 65535 CIGAR operations are stored (SAM spec section 4.2.2) with the placeholder
 * CIGAR `<l_qseq>S<ref_len>N` and the real operations in a CG:B,I tag.  htslib's bam_read1 moves them back
 * into place and drops the tag, so callers always see the real CIGAR; this does the same.
 * Returns 1 if the record was rewritten, 0 if it was left alone, -1 on error. */
static int restore_long_cigar(bam1_t *b) {
    bam1_core_t *c = &b->core;
    if (c->n_cigar == 0 || c->tid < 0 || c->pos < 0) return 0;
    uint32_t first;
    memcpy(&first, b->data + c->l_qname, 4);
    if (bam_cigar_op(first) != BAM_CSOFT_CLIP || (int32_t)bam_cigar_oplen(first) != c->l_qseq) return 0;
    const uint8_t *tag = bam_aux_get(b, "CG");
    if (!tag) return errno == ENOENT ? 0 : -1;
    if (tag[0] != 'B' || (tag[1] != 'I' && tag[1] != 'i')) return 0;
    const uint32_t n_real = ld32(tag + 2);
    if (n_real < c->n_cigar || n_real >= (1u << 29)) return 0;
    const size_t cig_at = c->l_qname, placeholder = (size_t)c->n_cigar * 4, real = (size_t)n_real * 4;
    const size_t tag_at = (size_t)(tag - b->data) - 2, tag_len = 8 + real;  /* "CG" 'B' 'I' count + payload */
    const size_t old_len = (size_t)b->l_data;
    if (tag_at + tag_len > old_len) return -1;
    /* new record = [.. cig_at) + real CIGAR + (cig_at + placeholder .. tag_at) + (tag_at + tag_len .. old_len) */
    uint8_t *ops = (uint8_t *)malloc(real ? real : 1);
    if (!ops) return -1;
    memcpy(ops, b->data + tag_at + 8, real);
    const size_t new_len = old_len - placeholder - 8;
    if (reserve(b, new_len + 8) != 0) { free(ops); return -1; }
    /* drop the tag first (it lies behind the CIGAR), then open the gap for the real operations */
    memmove(b->data + tag_at, b->data + tag_at + tag_len, old_len - tag_at - tag_len);
    const size_t after_cig = cig_at + placeholder, tail = old_len - tag_len - after_cig;
    memmove(b->data + cig_at + real, b->data + after_cig, tail);
    memcpy(b->data + cig_at, ops, real);
    free(ops);
    c->n_cigar = n_real;
    b->l_data = (int)new_len;
    return 1;
}

/* Returns bytes consumed (>=4) on success, -1 at EOF, < -1 on error. */
int bam_read1(BGZF *fp, bam1_t *b) {
    uint8_t x[36];
    ssize_t got = bgzf_read(fp, x, 4);
    if (got == 0) return -1;
    if (got != 4) return -3;
    int32_t block_len = (int32_t)ld32(x);
    if (block_len < 32) return -4;
    if (bgzf_read(fp, x + 4, 32) != 32) return -3;
    bam1_core_t *c = &b->core;
    c->tid = (int32_t)ld32(x + 4);
    c->pos = (int32_t)ld32(x + 8);
    uint32_t l_qname = x[12];
    c->qual = x[13];
    c->bin = ld16(x + 14);
    c->n_cigar = ld16(x + 16);
    c->flag = ld16(x + 18);
    c->l_qseq = (int32_t)ld32(x + 20);
    c->mtid = (int32_t)ld32(x + 24);
    c->mpos = (int32_t)ld32(x + 28);
    c->isize = (int32_t)ld32(x + 32);
    uint32_t extranul = (l_qname % 4) ? 4 - l_qname % 4 : 0;
    size_t payload = (size_t)block_len - 32;
    if (reserve(b, payload + extranul + 8) != 0) return -4;
    if (bgzf_read(fp, b->data, l_qname) != (ssize_t)l_qname) return -4;
    for (uint32_t i = 0; i < extranul; i++) b->data[l_qname + i] = 0;
    size_t rest = payload - l_qname;
    if (bgzf_read(fp, b->data + l_qname + extranul, rest) != (ssize_t)rest) return -4;
    c->l_extranul = (uint8_t)extranul;
    c->l_qname = (uint16_t)(l_qname + extranul);
    b->l_data = (int)(payload + extranul);
    size_t need = (size_t)c->l_qname + ((size_t)c->n_cigar << 2) + (((size_t)c->l_qseq + 1) >> 1) + (size_t)c->l_qseq;
    if (need > (size_t)b->l_data) return -4;
    if (restore_long_cigar(b) < 0) return -4;
    __atomic_fetch_add(&g_n_records, 1, __ATOMIC_RELAXED);
    __atomic_fetch_add(&g_n_bytes, (uint64_t)block_len + 4, __ATOMIC_RELAXED);
    return 4 + block_len;
}

int bam_write1(BGZF *fp, const bam1_t *b) {
    const bam1_core_t *c = &b->core;
    uint8_t x[36];
    uint32_t l_qname = c->l_qname - c->l_extranul;
    uint32_t block_len = (uint32_t)b->l_data - c->l_extranul + 32;
    const int long_cigar = c->n_cigar > 0xffffu;  /* placeholder CIGAR (8 bytes) + "CGBI" + count (8 bytes) */
    hts_pos_t ref_len = 0;
    if (long_cigar) {
        ref_len = bam_cigar2rlen((int)c->n_cigar, bam_get_cigar(b));
        if (ref_len >= (1 << 28)) return -1;  /* does not fit one CIGAR operation: not representable in BAM */
        block_len += 16;
    }
    st32(x, block_len);
    st32(x + 4, (uint32_t)c->tid);
    st32(x + 8, (uint32_t)c->pos);
    x[12] = (uint8_t)l_qname;
    x[13] = c->qual;
    st16(x + 14, c->bin);
    st16(x + 16, long_cigar ? 2 : (uint16_t)c->n_cigar);
    st16(x + 18, c->flag);
    st32(x + 20, (uint32_t)c->l_qseq);
    st32(x + 24, (uint32_t)c->mtid);
    st32(x + 28, (uint32_t)c->mpos);
    st32(x + 32, (uint32_t)c->isize);
    if (bgzf_flush_try(fp, 4 + block_len) != 0) return -1;
    if (bgzf_write(fp, x, 36) != 36) return -1;
    if (bgzf_write(fp, b->data, l_qname) != (ssize_t)l_qname) return -1;
    if (!long_cigar) {
        size_t rest = (size_t)b->l_data - c->l_qname;
        if (bgzf_write(fp, b->data + c->l_qname, rest) != (ssize_t)rest) return -1;
    } else {
        uint8_t ph[8], tg[8] = {'C', 'G', 'B', 'I', 0, 0, 0, 0};
        const size_t real = (size_t)c->n_cigar * 4;
        const uint8_t *after = b->data + c->l_qname + real;
        const size_t rest = (size_t)b->l_data - c->l_qname - real;
        st32(ph, (uint32_t)c->l_qseq << 4 | BAM_CSOFT_CLIP);
        st32(ph + 4, (uint32_t)ref_len << 4 | BAM_CREF_SKIP);
        st32(tg + 4, c->n_cigar);
        if (bgzf_write(fp, ph, 8) != 8) return -1;
        if (bgzf_write(fp, after, rest) != (ssize_t)rest) return -1;
        if (bgzf_write(fp, tg, 8) != 8) return -1;
        if (bgzf_write(fp, b->data + c->l_qname, real) != (ssize_t)real) return -1;
    }
    return (int)(4 + block_len);
}

hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar) {
    hts_pos_t l = 0;
    for (int k = 0; k < n_cigar; k++) {
        uint32_t v;
        memcpy(&v, cigar + k, 4);
        if (bam_cigar_type(bam_cigar_op(v)) & 2) l += bam_cigar_oplen(v);
    }
    return l;
}

int64_t bam_cigar2qlen(int n_cigar, const uint32_t *cigar) {
    int64_t l = 0;
    for (int k = 0; k < n_cigar; k++) {
        uint32_t v;
        memcpy(&v, cigar + k, 4);
        if (bam_cigar_type(bam_cigar_op(v)) & 1) l += bam_cigar_oplen(v);
    }
    return l;
}

hts_pos_t bam_endpos(const bam1_t *b) {
    hts_pos_t rlen = (b->core.flag & BAM_FUNMAP) ? 0 : bam_cigar2rlen((int)b->core.n_cigar, bam_get_cigar(b));
    if (rlen == 0) rlen = 1;
    return b->core.pos + rlen;
}

/* ---------- aux ---------- */

static int aux_type_size(int t) {
    switch (t) {
    case 'A': case 'c': case 'C': return 1;
    case 's': case 'S': return 2;
    case 'i': case 'I': case 'f': return 4;
    case 'd': return 8;
    default: return 0;
    }
}

/* s points at the type byte; returns pointer past the value, or NULL if malformed. */
static const uint8_t *aux_skip(const uint8_t *s, const uint8_t *end) {
    if (s >= end) return NULL;
    int t = *s++;
    int sz = aux_type_size(t);
    if (sz) return s + sz <= end ? s + sz : NULL;
    if (t == 'Z' || t == 'H') {
        while (s < end && *s) s++;
        return s < end ? s + 1 : NULL;
    }
    if (t == 'B') {
        if (s + 5 > end) return NULL;
        int esz = aux_type_size(*s);
        if (!esz) return NULL;
        uint32_t n = ld32(s + 1);
        s += 5;
        if ((uint64_t)n * esz > (uint64_t)(end - s)) return NULL;
        return s + (size_t)n * esz;
    }
    return NULL;
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
    const uint8_t *s = bam_get_aux(b);
    const uint8_t *end = b->data + b->l_data;
    while (s && s + 3 <= end) {
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return (uint8_t *)(s + 2);
        s = aux_skip(s + 2, end);
    }
    errno = s ? ENOENT : EINVAL;
    return NULL;
}

int64_t bam_aux2i(const uint8_t *s) {
    int t = *s++;
    switch (t) {
    case 'c': return (int8_t)s[0];
    case 'C': return s[0];
    case 's': return (int16_t)ld16(s);
    case 'S': return ld16(s);
    case 'i': return (int32_t)ld32(s);
    case 'I': return ld32(s);
    default: errno = EINVAL; return 0;
    }
}

double bam_aux2f(const uint8_t *s) {
    int t = *s;
    if (t == 'f') {
        float f;
        memcpy(&f, s + 1, 4);
        return f;
    }
    if (t == 'd') {
        double d;
        memcpy(&d, s + 1, 8);
        return d;
    }
    if (t == 'c' || t == 'C' || t == 's' || t == 'S' || t == 'i' || t == 'I') return (double)bam_aux2i(s);
    errno = EINVAL;
    return 0.0;
}

char *bam_aux2Z(const uint8_t *s) {
    if (*s == 'Z' || *s == 'H') return (char *)(s + 1);
    errno = EINVAL;
    return NULL;
}

int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data) {
    if (reserve(b, (size_t)b->l_data + 3 + len) != 0) return -1;
    uint8_t *p = b->data + b->l_data;
    p[0] = tag[0]; p[1] = tag[1]; p[2] = type;
    memcpy(p + 3, data, len);
    b->l_data += 3 + len;
    return 0;
}

/* Set integer tag to val, choosing the narrowest type for a new/enlarged tag
 * and keeping the slot width when the old slot is wide enough. */
int bam_aux_update_int(bam1_t *b, const char tag[2], int64_t val) {
    char type;
    uint32_t sz;
    if (val < INT32_MIN || val > UINT32_MAX) { errno = EOVERFLOW; return -1; }
    if (val < INT16_MIN) { type = 'i'; sz = 4; }
    else if (val < INT8_MIN) { type = 's'; sz = 2; }
    else if (val < 0) { type = 'c'; sz = 1; }
    else if (val < UINT8_MAX) { type = 'C'; sz = 1; }
    else if (val < UINT16_MAX) { type = 'S'; sz = 2; }
    else { type = 'I'; sz = 4; }

    uint8_t *s = bam_aux_get(b, tag);
    uint32_t old_sz = 0;
    int is_new = 0;
    if (s) {
        old_sz = (uint32_t)aux_type_size(*s);
        if (!(*s == 'c' || *s == 'C' || *s == 's' || *s == 'S' || *s == 'i' || *s == 'I')) { errno = EINVAL; return -1; }
    } else {
        if (errno != ENOENT) return -1;
        is_new = 1;
    }
    if (is_new || old_sz < sz) {
        size_t s_off = is_new ? (size_t)b->l_data : (size_t)(s - b->data);
        size_t grow = (is_new ? 3 : 0) + sz - old_sz;
        if (reserve(b, (size_t)b->l_data + grow) != 0) return -1;
        s = b->data + s_off;
        if (is_new) {
            *s++ = tag[0];
            *s++ = tag[1];
        } else {
            memmove(s + 1 + sz, s + 1 + old_sz, (size_t)b->l_data - s_off - 1 - old_sz);
        }
        b->l_data += (int)grow;
    } else {
        sz = old_sz;
        type = (val < 0 ? "\0cs\0i" : "\0CS\0I")[old_sz];
    }
    *s++ = (uint8_t)type;
    uint32_t u = (uint32_t)val;
    for (uint32_t i = 0; i < sz; i++) s[i] = (uint8_t)(u >> (8 * i));
    return 0;
}
