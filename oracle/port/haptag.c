/* TEST INFRASTRUCTURE (oracle port) — read haplotagging for untagged BAMs (-u).
 *
 * Restates parse_variants_for_one_read (reference blockjoin.c:1545-1691) and
 * haptag_one_read_with_variants (blockjoin.c:1693-1840) on a packed record and a flat known-variant
 * array (the output of insert_variant_from_vcf_line, blockjoin.c:1432-1543).
 */
#include <stdlib.h>
#include <string.h>
#include "port.h"

typedef struct { uint32_t pos, len; uint8_t op; uint8_t *bases; } rvar_t;
typedef struct { rvar_t *a; size_t n, m; } rvars_t;

static inline int nib(const uint8_t *seq, uint32_t i) { return (seq[i >> 1] >> ((~i & 1) << 2)) & 0xf; }
/* seq_nt4_table[seq_nt16_str[nibble]]: A0 C1 G2 T3 everything else 4 (blockjoin.c:74-92) */
static inline uint8_t nt4_of_nib(int c) { return c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4; }
static inline uint8_t nt4_of_char(int ch) {
    switch (ch) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3; default: return 4;
    }
}
/* md_op_table (blockjoin.c:94-115): digits 0, '^' 1, ACGTUN (either case) 2, else 4 */
static inline int md_class(int ch) {
    if (ch >= '0' && ch <= '9') return 0;
    if (ch == '^') return 1;
    switch (ch) {
    case 'A': case 'C': case 'G': case 'T': case 'U': case 'N':
    case 'a': case 'c': case 'g': case 't': case 'u': case 'n': return 2;
    default: return 4;
    }
}

static rvar_t *rv_push(rvars_t *v, uint32_t pos, uint8_t op, uint32_t len) {
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 32; v->a = (rvar_t *)realloc(v->a, sizeof(rvar_t) * v->m); }
    rvar_t *r = &v->a[v->n++];
    r->pos = pos; r->op = op; r->len = len;
    r->bases = (uint8_t *)malloc(len ? len : 1);
    return r;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

static int parse_read_variants(const pomfret_gpu_read_desc *r, rvars_t *out) {
    /* pass 1: CIGAR, :1564-1589 */
    uint64_t *ins = (uint64_t *)malloc(sizeof(uint64_t) * (r->n_cigar ? r->n_cigar : 1));
    size_t n_ins = 0;
    uint32_t self_start = 0, ref_pos = r->pos, self_pos = 0;
    for (uint32_t i = 0; i < r->n_cigar; i++) {
        uint32_t op = r->cigar[i] & 15, L = r->cigar[i] >> 4;
        if (op == 3) ref_pos += L;
        else if (op == 4) { if (i == 0) self_start = L; self_pos += L; }
        else if (op == 0 || op == 7 || op == 8) { ref_pos += L; self_pos += L; }
        else if (op == 1) {
            rvar_t *v = rv_push(out, ref_pos, 2 /*VAR_OP_I*/, L);
            for (uint32_t j = 0; j < L; j++) v->bases[j] = nt4_of_nib(nib(r->seq, self_pos + j));
            ins[n_ins++] = ((uint64_t)L) << 32 | self_pos;
            self_pos += L;
        } else if (op == 2) ref_pos += L;
    }
    /* pass 2: MD, :1591-1673 */
    if (!r->md) { free(ins); return POMFRET_GPU_ERR_MISSING_MD; }
    const char *md = r->md;
    const uint32_t n_md = r->md_len;
    size_t prev_ins = 0;
    self_pos = self_start;
    ref_pos = r->pos;
    int prev_type = n_md ? md_class((unsigned char)md[0]) : 4;
    uint32_t prev_i = 0;
    if (n_md == 0) { free(ins); return 0; } /* md_op_table['\0'] = 4 would trip the assert; treat "" as no-op */
    if (prev_type == 2) {
        rvar_t *v = rv_push(out, ref_pos, 1 /*VAR_OP_X*/, 1);
        v->bases[0] = nt4_of_nib(nib(r->seq, self_pos));
        ref_pos++; self_pos++;
        prev_type = -1;
    }
    if (prev_type >= 4) { free(ins); return POMFRET_GPU_ERR_BAD_MD; }
    for (uint32_t i = 1; i < n_md; i++) {
        int t = md_class((unsigned char)md[i]);
        if (t == 4) { free(ins); return POMFRET_GPU_ERR_BAD_MD; }
        if (t == prev_type) continue;
        if (prev_type == 0) {
            int l = 0;
            for (uint32_t j = prev_i; j < i; j++) l = l * 10 + (md[j] - '0'); /* natoi, :117-130 */
            ref_pos += (uint32_t)l;
            self_pos += (uint32_t)l;
            while (prev_ins < n_ins && self_pos > (uint32_t)ins[prev_ins]) {
                self_pos += (uint32_t)(ins[prev_ins] >> 32);
                prev_ins++;
            }
        } else if (prev_type == 1) {
            if (t == 0) {
                uint32_t L = i - prev_i - 1;
                rvar_t *v = rv_push(out, ref_pos, 3 /*VAR_OP_D*/, L);
                for (uint32_t j = 0; j < L; j++) v->bases[j] = nt4_of_char((unsigned char)md[prev_i + 1 + j]);
                ref_pos += L;
                prev_type = t;
                prev_i = i;
            }
            continue;
        }
        if (t == 2) {
            rvar_t *v = rv_push(out, ref_pos, 1, 1);
            v->bases[0] = nt4_of_nib(nib(r->seq, self_pos));
            ref_pos++; self_pos++;
            prev_type = -1;
            prev_i = i;
        } else {
            prev_type = t;
            prev_i = i;
        }
    }
    free(ins);
    return 0;
}

void port_haptag_cursors(const uint32_t *start_pos, int n_reads, const pomfret_gpu_variant *known, uint32_t n_known,
                         uint32_t *known_first) {
    uint32_t prev = 0;
    for (int r = 0; r < n_reads; r++) {
        uint32_t i = prev;
        if (n_known == 0) { known_first[r] = 0; continue; }
        while (i < n_known && known[i].pos < start_pos[r]) i++; /* uint32 vs int start_pos: unsigned compare */
        prev = i == 0 ? 0 : i - 1;
        known_first[r] = i;
    }
}

int port_haptag_read(const pomfret_gpu_read_desc *r, const pomfret_gpu_variant *known, uint32_t n_known,
                     const uint8_t *bases, uint32_t known_first, int *votes) {
    rvars_t rv = {0, 0, 0};
    int rc = parse_read_variants(r, &rv);
    int tag = 254;
    int cnt[2] = {0, 0};
    if (rc == 0 && n_known > 0) {
        uint64_t rlen = 0;
        for (uint32_t i = 0; i < r->n_cigar; i++) {
            uint32_t op = r->cigar[i] & 15;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += r->cigar[i] >> 4;
        }
        if (rlen == 0) rlen = 1;
        const int end_pos = (int)(r->pos + rlen);
        size_t cap = rv.n + 16, n = 0;
        uint64_t *pb = (uint64_t *)malloc(sizeof(uint64_t) * cap);
        const uint64_t typebit = 1ull << 32;
        for (uint32_t i = known_first; i < n_known; i++) {
            if (known[i].pos >= (uint32_t)end_pos) break;
            if (n == cap) { cap *= 2; pb = (uint64_t *)realloc(pb, sizeof(uint64_t) * cap); }
            pb[n++] = ((uint64_t)known[i].pos) << 33 | i;
        }
        for (uint32_t i = 0; i < rv.n; i++) {
            if (n == cap) { cap *= 2; pb = (uint64_t *)realloc(pb, sizeof(uint64_t) * cap); }
            pb[n++] = ((uint64_t)rv.a[i].pos) << 33 | typebit | i;
        }
        qsort(pb, n, sizeof(uint64_t), cmp_u64);
        for (size_t i = 0; i < n;) { /* :1749-1814 */
            if (pb[i] & typebit) { i++; continue; }
            uint32_t ref_pos = (uint32_t)(pb[i] >> 33), ref_i = (uint32_t)pb[i];
            if (i + 1 == n) { cnt[known[ref_i].haptag]++; break; }
            uint32_t self_pos = (uint32_t)(pb[i + 1] >> 33), self_i = (uint32_t)pb[i + 1];
            if (ref_pos != self_pos) {
                int skip = 0;
                if (i > 0 && (pb[i - 1] & typebit)) {
                    uint32_t lp = (uint32_t)(pb[i - 1] >> 33), li = (uint32_t)pb[i - 1];
                    if (rv.a[li].op == 3 && lp + rv.a[li].len >= ref_pos) skip = 1;
                }
                if (!skip) cnt[known[ref_i].haptag]++;
                i++;
            } else {
                if (!(pb[i + 1] & typebit)) { i += 2; continue; } /* two known variants on one position */
                const pomfret_gpu_variant *kv = &known[ref_i];
                const rvar_t *s = &rv.a[self_i];
                int ok = kv->len == s->len;
                if (ok) for (uint32_t j = 0; j < kv->len; j++) if (bases[kv->bases_off + j] != s->bases[j]) { ok = 0; break; }
                if (ok) cnt[kv->haptag ^ 1]++;
                i += 2;
            }
        }
        free(pb);
        /* :1817-1832 */
        float mx = (float)(cnt[0] > cnt[1] ? cnt[0] : cnt[1]);
        int mn = cnt[0] <= cnt[1] ? cnt[0] : cnt[1];
        float ratio = mn == 0 ? 0 : mx / (float)mn;
        if ((cnt[0] > 3 && cnt[1] > 3 && ratio < 5) || cnt[0] == cnt[1]) tag = 254;
        else tag = cnt[0] > cnt[1] ? 0 : 1;
    }
    if (votes) { votes[0] = cnt[0]; votes[1] = cnt[1]; }
    for (size_t i = 0; i < rv.n; i++) free(rv.a[i].bases);
    free(rv.a);
    return rc < 0 ? rc : tag;
}
