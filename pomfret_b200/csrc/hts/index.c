/* hts-shim BAI index: load, region iterator, build (SAMv1 §5.2; binning
 * scheme §5.3 with min_shift=14, depth=5).  Reference call sites:
 * blockjoin.c:579 (load), :1061/:1853/:2512/:3032 (query), :4723 (build). */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "htslib/sam.h"

typedef struct { uint64_t beg, end; } chunk_t;
typedef struct { uint32_t bin; int n; chunk_t *chunks; } bin_t;
typedef struct { int n_bin; bin_t *bins; int n_lin; uint64_t *lin; } ref_index_t;

struct hts_idx_t {
    int n_ref;
    ref_index_t *refs;
};

struct hts_itr_t {
    int whole_file; /* "." : every record from the first one */
    int started;
    int finished;
    int tid;
    hts_pos_t beg, end;
    int n_chunks, i_chunk;
    chunk_t *chunks;
    uint64_t cur;
};

static inline uint32_t ld32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint64_t ld64(const uint8_t *p) { return (uint64_t)ld32(p) | ((uint64_t)ld32(p + 4) << 32); }

void hts_idx_destroy(hts_idx_t *idx) {
    if (!idx) return;
    for (int i = 0; i < idx->n_ref; i++) {
        for (int j = 0; j < idx->refs[i].n_bin; j++) free(idx->refs[i].bins[j].chunks);
        free(idx->refs[i].bins);
        free(idx->refs[i].lin);
    }
    free(idx->refs);
    free(idx);
}

void hts_itr_destroy(hts_itr_t *itr) {
    if (!itr) return;
    free(itr->chunks);
    free(itr);
}

static uint8_t *slurp(const char *fn, size_t *len) {
    FILE *f = fopen(fn, "rb");
    if (!f) return NULL;
    fseeko(f, 0, SEEK_END);
    off_t n = ftello(f);
    fseeko(f, 0, SEEK_SET);
    uint8_t *buf = (uint8_t *)malloc(n > 0 ? (size_t)n : 1);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); fclose(f); return NULL; }
    fclose(f);
    *len = (size_t)n;
    return buf;
}

hts_idx_t *sam_index_load(htsFile *fp, const char *fn) {
    (void)fp;
    size_t l = strlen(fn), len = 0;
    char *fnidx = (char *)malloc(l + 8);
    sprintf(fnidx, "%s.bai", fn);
    uint8_t *buf = slurp(fnidx, &len);
    if (!buf && l > 4 && strcmp(fn + l - 4, ".bam") == 0) {
        sprintf(fnidx, "%.*s.bai", (int)(l - 4), fn);
        buf = slurp(fnidx, &len);
    }
    free(fnidx);
    if (!buf) return NULL;
    if (len < 8 || memcmp(buf, "BAI\1", 4) != 0) { free(buf); return NULL; }
    hts_idx_t *idx = (hts_idx_t *)calloc(1, sizeof(hts_idx_t));
    const uint8_t *p = buf + 4, *end = buf + len;
    idx->n_ref = (int)ld32(p); p += 4;
    idx->refs = (ref_index_t *)calloc(idx->n_ref > 0 ? idx->n_ref : 1, sizeof(ref_index_t));
    for (int i = 0; i < idx->n_ref; i++) {
        ref_index_t *r = &idx->refs[i];
        if (p + 4 > end) goto bad;
        r->n_bin = (int)ld32(p); p += 4;
        r->bins = (bin_t *)calloc(r->n_bin > 0 ? r->n_bin : 1, sizeof(bin_t));
        for (int j = 0; j < r->n_bin; j++) {
            if (p + 8 > end) goto bad;
            r->bins[j].bin = ld32(p);
            r->bins[j].n = (int)ld32(p + 4);
            p += 8;
            if (p + 16 * (size_t)r->bins[j].n > end) goto bad;
            r->bins[j].chunks = (chunk_t *)malloc(sizeof(chunk_t) * (r->bins[j].n > 0 ? r->bins[j].n : 1));
            for (int k = 0; k < r->bins[j].n; k++) {
                r->bins[j].chunks[k].beg = ld64(p);
                r->bins[j].chunks[k].end = ld64(p + 8);
                p += 16;
            }
        }
        if (p + 4 > end) goto bad;
        r->n_lin = (int)ld32(p); p += 4;
        if (p + 8 * (size_t)r->n_lin > end) goto bad;
        r->lin = (uint64_t *)malloc(sizeof(uint64_t) * (r->n_lin > 0 ? r->n_lin : 1));
        for (int k = 0; k < r->n_lin; k++) { r->lin[k] = ld64(p); p += 8; }
    }
    free(buf);
    return idx;
bad:
    free(buf);
    hts_idx_destroy(idx);
    return NULL;
}

/* bins overlapping [beg,end) — SAMv1 §5.3 */
static int reg2bins(int64_t beg, int64_t end, uint16_t *list) {
    int i = 0, k;
    --end;
    list[i++] = 0;
    for (k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) list[i++] = k;
    for (k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) list[i++] = k;
    for (k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) list[i++] = k;
    for (k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) list[i++] = k;
    for (k = 4681 + (beg >> 14); k <= 4681 + (end >> 14); ++k) list[i++] = k;
    return i;
}

static int cmp_chunk(const void *a, const void *b) {
    const chunk_t *x = (const chunk_t *)a, *y = (const chunk_t *)b;
    return x->beg < y->beg ? -1 : x->beg > y->beg ? 1 : 0;
}

hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end) {
    hts_itr_t *it = (hts_itr_t *)calloc(1, sizeof(hts_itr_t));
    if (tid < 0) { /* whole file */
        it->whole_file = 1;
        return it;
    }
    it->tid = tid;
    if (beg < 0) beg = 0;
    if (end > ((hts_pos_t)1 << 29)) end = (hts_pos_t)1 << 29;
    it->beg = beg;
    it->end = end;
    if (!idx || tid >= idx->n_ref || beg >= end) { it->finished = 1; return it; }
    const ref_index_t *r = &idx->refs[tid];
    uint64_t min_off = 0;
    if (r->n_lin > 0) {
        int w = (int)(beg >> 14);
        min_off = w < r->n_lin ? r->lin[w] : r->lin[r->n_lin - 1];
        if (w < r->n_lin && min_off == 0) { /* empty window: take the previous non-empty one */
            for (int k = w; k >= 0; k--) if (r->lin[k]) { min_off = r->lin[k]; break; }
        }
    }
    uint16_t *bins = (uint16_t *)malloc(sizeof(uint16_t) * 37450);
    int nb = reg2bins(beg, end, bins);
    int cap = 16, n = 0;
    chunk_t *cs = (chunk_t *)malloc(sizeof(chunk_t) * cap);
    for (int j = 0; j < r->n_bin; j++) {
        uint32_t b = r->bins[j].bin;
        if (b >= 37450) continue; /* metadata pseudo-bin */
        int hit = 0;
        for (int k = 0; k < nb; k++) if (bins[k] == b) { hit = 1; break; }
        if (!hit) continue;
        for (int k = 0; k < r->bins[j].n; k++) {
            if (r->bins[j].chunks[k].end <= min_off) continue;
            if (n == cap) { cap *= 2; cs = (chunk_t *)realloc(cs, sizeof(chunk_t) * cap); }
            cs[n++] = r->bins[j].chunks[k];
        }
    }
    free(bins);
    qsort(cs, n, sizeof(chunk_t), cmp_chunk);
    int m = 0;
    for (int k = 0; k < n; k++) {
        if (m > 0 && cs[k].beg <= cs[m - 1].end) {
            if (cs[k].end > cs[m - 1].end) cs[m - 1].end = cs[k].end;
        } else cs[m++] = cs[k];
    }
    it->chunks = cs;
    it->n_chunks = m;
    if (m == 0) it->finished = 1;
    return it;
}

static int64_t parse_num(const char *s, const char *e, int *ok) {
    int64_t v = 0;
    int nd = 0;
    for (; s < e; s++) {
        if (*s == ',') continue;
        if (*s < '0' || *s > '9') { *ok = 0; return 0; }
        v = v * 10 + (*s - '0');
        nd++;
    }
    *ok = nd > 0;
    return v;
}

hts_itr_t *sam_itr_querys(const hts_idx_t *idx, sam_hdr_t *hdr, const char *region) {
    if (!hdr || !region) return NULL;
    if (strcmp(region, ".") == 0) return sam_itr_queryi(idx, -1, 0, 0);
    if (!idx) return NULL;
    int tid = sam_hdr_name2tid(hdr, region);
    hts_pos_t beg = 0, end = (hts_pos_t)1 << 29;
    if (tid < 0) {
        const char *colon = strrchr(region, ':');
        if (!colon) return NULL;
        size_t nl = (size_t)(colon - region);
        char *name = (char *)malloc(nl + 1);
        memcpy(name, region, nl);
        name[nl] = 0;
        tid = sam_hdr_name2tid(hdr, name);
        free(name);
        if (tid < 0) return NULL;
        const char *s = colon + 1, *e = region + strlen(region);
        const char *dash = strchr(s, '-');
        int ok = 1;
        if (dash && dash > s) {
            beg = parse_num(s, dash, &ok) - 1;
            if (!ok) return NULL;
            if (dash + 1 < e) {
                end = parse_num(dash + 1, e, &ok);
                if (!ok) return NULL;
            }
        } else if (!dash) {
            beg = parse_num(s, e, &ok) - 1;
            if (!ok) return NULL;
        } else return NULL;
        if (beg < 0) beg = 0;
    }
    return sam_itr_queryi(idx, tid, beg, end);
}

/* shim extension: the file chunks an iterator would walk, as (begin, end) virtual offsets in file order
 * (a loader that ships the BGZF blocks to the device instead of inflating them here) */
/* linear index of one target: virtual offsets of the first record that overlaps each 16 kb window (record starts) */
int pomfret_idx_linear(const hts_idx_t *idx, int tid, const uint64_t **lin) {
    if (!idx || tid < 0 || tid >= idx->n_ref) return 0;
    *lin = idx->refs[tid].lin;
    return idx->refs[tid].n_lin;
}
int pomfret_idx_nref(const hts_idx_t *idx) { return idx ? idx->n_ref : 0; }

int pomfret_itr_chunks(const hts_itr_t *itr, const uint64_t **pairs) {
    if (!itr || itr->whole_file) return -1;
    *pairs = (const uint64_t *)itr->chunks;
    return itr->n_chunks;
}

int sam_itr_next(htsFile *htsfp, hts_itr_t *itr, bam1_t *r) {
    if (!itr || !htsfp) return -2;
    BGZF *fp = htsfp->fp.bgzf;
    if (itr->finished) return -1;
    if (itr->whole_file) {
        if (!itr->started) {
            if (bgzf_seek(fp, 0, SEEK_SET) != 0) return -2;
            sam_hdr_t *h = bam_hdr_read(fp);
            if (!h) return -2;
            sam_hdr_destroy(h);
            itr->started = 1;
        }
        int rc = bam_read1(fp, r);
        if (rc < 0) { itr->finished = 1; return rc < -1 ? -2 : -1; }
        return rc;
    }
    for (;;) {
        if (!itr->started || itr->cur >= itr->chunks[itr->i_chunk].end) {
            if (itr->started) itr->i_chunk++;
            if (itr->i_chunk >= itr->n_chunks) { itr->finished = 1; return -1; }
            if (!itr->started || itr->cur < itr->chunks[itr->i_chunk].beg) {
                if (bgzf_seek(fp, (int64_t)itr->chunks[itr->i_chunk].beg, SEEK_SET) != 0) return -2;
                itr->cur = itr->chunks[itr->i_chunk].beg;
            }
            itr->started = 1;
        }
        int rc = bam_read1(fp, r);
        if (rc < 0) { itr->finished = 1; return rc < -1 ? -2 : -1; }
        itr->cur = (uint64_t)bgzf_tell(fp);
        if (r->core.tid != itr->tid || r->core.pos >= itr->end) { itr->finished = 1; return -1; }
        if (bam_endpos(r) > itr->beg) return rc;
    }
}

/* ---------- index build ---------- */

typedef struct { uint32_t bin; int n, m; chunk_t *a; } bbin_t;
typedef struct {
    int n_bin, m_bin;
    bbin_t *bins;
    int n_lin, m_lin;
    uint64_t *lin;
    uint64_t off_beg, off_end, n_mapped, n_unmapped;
    int seen;
} bref_t;

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

static bbin_t *get_bin(bref_t *r, uint32_t bin) {
    /* records are coordinate sorted, so the most recent bins are the likely hits */
    for (int i = r->n_bin - 1; i >= 0; i--)
        if (r->bins[i].bin == bin) return &r->bins[i];
    if (r->n_bin == r->m_bin) {
        r->m_bin = r->m_bin ? r->m_bin * 2 : 64;
        r->bins = (bbin_t *)realloc(r->bins, sizeof(bbin_t) * r->m_bin);
    }
    bbin_t *b = &r->bins[r->n_bin++];
    memset(b, 0, sizeof(*b));
    b->bin = bin;
    return b;
}

static int cmp_bbin(const void *a, const void *b) {
    uint32_t x = ((const bbin_t *)a)->bin, y = ((const bbin_t *)b)->bin;
    return x < y ? -1 : x > y ? 1 : 0;
}

static void w32(FILE *f, uint32_t v) {
    uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
    fwrite(b, 1, 4, f);
}
static void w64(FILE *f, uint64_t v) { w32(f, (uint32_t)v); w32(f, (uint32_t)(v >> 32)); }

/* Incremental builder: one call per record in file order with the virtual offsets in front of and behind it.
 * sam_index_build3() feeds it from a finished file; a writer that knows its own offsets (the synthetic data
 * generator) feeds it directly and saves the second pass over the file. */
struct pomfret_bai_builder {
    int n_ref;
    bref_t *refs;
    uint64_t n_no_coor;
    int last_tid;
    hts_pos_t last_pos;
};

pomfret_bai_builder *pomfret_bai_new(int n_ref) {
    pomfret_bai_builder *bb = (pomfret_bai_builder *)calloc(1, sizeof(*bb));
    bb->n_ref = n_ref;
    bb->refs = (bref_t *)calloc(n_ref > 0 ? n_ref : 1, sizeof(bref_t));
    bb->last_tid = -1;
    bb->last_pos = -1;
    return bb;
}

int pomfret_bai_add(pomfret_bai_builder *B, int tid, hts_pos_t beg, hts_pos_t end, int unmapped, uint64_t off0, uint64_t off1) {
    if (tid < 0) { B->n_no_coor++; return 0; }
    if (tid >= B->n_ref || tid < B->last_tid || (tid == B->last_tid && beg < B->last_pos)) return -1;
    B->last_tid = tid;
    B->last_pos = beg;
    bref_t *r = &B->refs[tid];
    if (!r->seen) { r->seen = 1; r->off_beg = off0; }
    r->off_end = off1;
    if (unmapped) r->n_unmapped++; else r->n_mapped++;
    bbin_t *bb = get_bin(r, (uint32_t)reg2bin(beg, end));
    if (bb->n > 0 && bb->a[bb->n - 1].end == off0) bb->a[bb->n - 1].end = off1;
    else {
        if (bb->n == bb->m) { bb->m = bb->m ? bb->m * 2 : 4; bb->a = (chunk_t *)realloc(bb->a, sizeof(chunk_t) * bb->m); }
        bb->a[bb->n].beg = off0;
        bb->a[bb->n].end = off1;
        bb->n++;
    }
    int w0 = (int)(beg >> 14), w1 = (int)((end - 1) >> 14);
    if (w1 + 1 > r->m_lin) {
        int m = r->m_lin ? r->m_lin : 64;
        while (m < w1 + 1) m *= 2;
        r->lin = (uint64_t *)realloc(r->lin, sizeof(uint64_t) * m);
        memset(r->lin + r->m_lin, 0, sizeof(uint64_t) * (m - r->m_lin));
        r->m_lin = m;
    }
    for (int w = w0; w <= w1; w++) if (r->lin[w] == 0) r->lin[w] = off0;
    if (w1 + 1 > r->n_lin) r->n_lin = w1 + 1;
    return 0;
}

/* writes the index (fnidx may be NULL: just discard) and frees the builder */
int pomfret_bai_finish(pomfret_bai_builder *B, const char *fnidx) {
    int ret = 0;
    bref_t *refs = B->refs;
    const int n_ref = B->n_ref;
    if (fnidx) {
        FILE *f = fopen(fnidx, "wb");
        if (!f) ret = -4;
        else {
            fwrite("BAI\1", 1, 4, f);
            w32(f, (uint32_t)n_ref);
            for (int i = 0; i < n_ref; i++) {
                bref_t *r = &refs[i];
                qsort(r->bins, r->n_bin, sizeof(bbin_t), cmp_bbin);
                /* fill empty linear windows with the following non-empty offset's predecessor */
                for (int w = 1; w < r->n_lin; w++) if (r->lin[w] == 0) r->lin[w] = r->lin[w - 1];
                w32(f, (uint32_t)(r->n_bin + (r->seen ? 1 : 0)));
                for (int j = 0; j < r->n_bin; j++) {
                    w32(f, r->bins[j].bin);
                    w32(f, (uint32_t)r->bins[j].n);
                    for (int k = 0; k < r->bins[j].n; k++) { w64(f, r->bins[j].a[k].beg); w64(f, r->bins[j].a[k].end); }
                }
                if (r->seen) {
                    w32(f, 37450);
                    w32(f, 2);
                    w64(f, r->off_beg); w64(f, r->off_end);
                    w64(f, r->n_mapped); w64(f, r->n_unmapped);
                }
                w32(f, (uint32_t)r->n_lin);
                for (int w = 0; w < r->n_lin; w++) w64(f, r->lin[w]);
            }
            w64(f, B->n_no_coor);
            if (fclose(f) != 0) ret = -4;
        }
    }
    for (int i = 0; i < n_ref; i++) {
        for (int j = 0; j < refs[i].n_bin; j++) free(refs[i].bins[j].a);
        free(refs[i].bins);
        free(refs[i].lin);
    }
    free(refs);
    free(B);
    return ret;
}

int sam_index_build3(const char *fn, const char *fnidx, int min_shift, int nthreads) {
    (void)nthreads;
    if (min_shift != 0) return -1; /* BAI only */
    htsFile *fp = hts_open(fn, "r");
    if (!fp) return -2;
    sam_hdr_t *h = sam_hdr_read(fp);
    if (!h) { hts_close(fp); return -1; }
    pomfret_bai_builder *B = pomfret_bai_new(h->n_targets);
    bam1_t *b = bam_init1();
    BGZF *bg = fp->fp.bgzf;
    uint64_t off0 = (uint64_t)bgzf_tell(bg);
    int rc, ret = 0;
    while ((rc = bam_read1(bg, b)) >= 0) {
        uint64_t off1 = (uint64_t)bgzf_tell(bg);
        if (pomfret_bai_add(B, b->core.tid, b->core.pos, bam_endpos(b), (b->core.flag & BAM_FUNMAP) != 0, off0, off1) != 0) { ret = -1; break; }
        off0 = off1;
    }
    if (rc < -1) ret = -1;
    bam_destroy1(b);
    int rs = pomfret_bai_finish(B, ret == 0 ? fnidx : NULL);
    if (ret == 0) ret = rs;
    sam_hdr_destroy(h);
    hts_close(fp);
    return ret;
}
