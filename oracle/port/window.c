/* TEST INFRASTRUCTURE (oracle port) — the per-window engine.
 *
 * Restates haplotag_region_given_bam (reference blockjoin.c:4217-4335) and everything below it on
 * flat arrays: read set construction (1043-1173), site pileup + methmer layout (3202-3354), methmer
 * extraction (3357-3451, 339-421), count tables and scoring (3453-3515, 3576-3656), range growth
 * (3669-3691), greedy propagation (3693-3774, 3958-4080), evaluation (3881-3956) and the join rule
 * (4088-4214 with one permutation, 4313-4320).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "port.h"

/* ---------------- sites ---------------- */

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

typedef struct { uint32_t pos; uint16_t cnt[3]; int used; } pile_t;

void port_sites(const port_readset_t *rs, int cov_sel, int k, int k_span, int direction, port_sites_t *out) {
    /* blockjoin.c:3210-3253: per position u16 counters, count in bits 4.., strand saturating bits below.
     * Open addressing table instead of khashl; iteration order is irrelevant (sites are sorted, :3298). */
    size_t tot = 0;
    for (int i = 0; i < rs->n; i++) tot += rs->calls[i].n;
    size_t cap = 64;
    while (cap < tot * 2 + 8) cap <<= 1;
    pile_t *ht = (pile_t *)calloc(cap, sizeof(pile_t));
    for (int i = 0; i < rs->n; i++) {
        uint8_t strand = rs->strand[i];
        for (uint32_t j = 0; j < rs->calls[i].n; j++) {
            uint32_t pos = rs->calls[i].pos[j];
            int call = rs->calls[i].cat[j];
            size_t h = (pos * 2654435761u) & (cap - 1);
            while (ht[h].used && ht[h].pos != pos) h = (h + 1) & (cap - 1);
            if (!ht[h].used) {
                ht[h].used = 1; ht[h].pos = pos;
                ht[h].cnt[0] = ht[h].cnt[1] = ht[h].cnt[2] = 0;
                ht[h].cnt[call] = 16;
            } else {
                ht[h].cnt[call] += 16; /* the reference's limit test is always true; u16 wraps */
            }
            if (strand == 0) { if ((ht[h].cnt[call] & 3) < 3) ht[h].cnt[call] += 1; }
            else { if (((ht[h].cnt[call] >> 2) & 3) < 3) ht[h].cnt[call] += 4; }
        }
    }
    int n = 0;
    for (size_t h = 0; h < cap; h++)
        if (ht[h].used && (ht[h].cnt[0] >> 4) >= cov_sel && (ht[h].cnt[1] >> 4) >= cov_sel) n++;
    out->n = n;
    out->real_pos = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    out->starts = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    out->lens = (uint8_t *)calloc(n ? n : 1, 1);
    n = 0;
    for (size_t h = 0; h < cap; h++)
        if (ht[h].used && (ht[h].cnt[0] >> 4) >= cov_sel && (ht[h].cnt[1] >> 4) >= cov_sel) out->real_pos[n++] = ht[h].pos;
    free(ht);
    qsort(out->real_pos, n, sizeof(uint32_t), cmp_u32);
    uint32_t *sites = out->real_pos;
    /* :3299-3329 — for direction 1 run the same loop on the reversed array, then reverse everything back */
#define REV32(a, l) for (int x_ = 0; x_ < (l) / 2; x_++) { uint32_t t_ = (a)[x_]; (a)[x_] = (a)[(l)-1-x_]; (a)[(l)-1-x_] = t_; }
#define REV8(a, l) for (int x_ = 0; x_ < (l) / 2; x_++) { uint8_t t_ = (a)[x_]; (a)[x_] = (a)[(l)-1-x_]; (a)[(l)-1-x_] = t_; }
    if (direction == 1) REV32(sites, n);
    for (int i = 0; i < n; i++) {
        int j = i + k;
        if (j > n - 1) j = n - 1;
        for (;;) {
            if (direction == 0 && (uint32_t)(sites[j] - sites[i]) <= (uint32_t)k_span) break;
            if (direction == 1 && (uint32_t)(sites[i] - sites[j]) <= (uint32_t)k_span) break;
            j--;
        }
        out->lens[i] = (uint8_t)(j - i == 0 ? 1 : j - i);
        out->starts[i] = direction == 0 ? sites[i] : sites[j];
    }
    if (direction == 1) { REV32(sites, n); REV32(out->starts, n); REV8(out->lens, n); }
}

void port_sites_free(port_sites_t *s) {
    free(s->real_pos); free(s->starts); free(s->lens);
    s->real_pos = s->starts = NULL; s->lens = NULL; s->n = 0;
}

/* ---------------- methmers of one read ---------------- */

/* search_arr1 + search_arr(which_end=0), blockjoin.c:339-421 */
static int search_left(const uint32_t *a, uint32_t l, uint32_t v, uint32_t *idx) {
    if (l == 0) return -3;
    if (v < a[0]) { *idx = UINT32_MAX; return -1; }
    if (v > a[l - 1]) { *idx = UINT32_MAX; return -2; }
    uint32_t i = 0;
    int stat = 0;
    if (l < 16) {
        for (i = 0; i < l; i++) {
            if (a[i] == v) { stat = 1; break; }
            if (a[i] > v) { stat = 0; break; }
        }
    } else {
        uint32_t lo = 0, hi = l - 1;
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo) / 2;
            if (v <= a[mid]) hi = mid; else lo = mid + 1;
        }
        i = hi;
        stat = a[hi] == v;
    }
    if (stat == 1) while (i > 0 && a[i - 1] == v) i--;
    *idx = i;
    return stat;
}

int port_mmr_of_read(const port_calls_t *calls, const port_sites_t *ms, uint32_t **out, uint32_t *start_i) {
    *out = NULL;
    *start_i = UINT32_MAX;
    const uint32_t *sites = ms->starts; /* searches run on sites_starts, :3363 */
    const uint32_t n_sites = (uint32_t)ms->n;
    const uint32_t nc = calls->n;
    if (nc == 0) return 0;
    uint32_t xl = 0, xr = 0;
    int stat = search_left(sites, n_sites, calls->pos[0], &xl);
    if (stat == -2 || stat == -3) return 0;
    if (stat == 0) xl = xl == 0 ? 0 : xl - 1;
    stat = search_left(sites, n_sites, calls->pos[nc - 1], &xr);
    if (stat == -1 || stat == -3) return 0;
    if (xl == UINT32_MAX) xl = 0;
    if (xr == UINT32_MAX) xr = n_sites;

    size_t cap = (xr > xl ? xr - xl : 0) + nc + 1, nb = 0;
    uint64_t *buf = (uint64_t *)malloc(sizeof(uint64_t) * cap);
    for (uint32_t i = xl; i < xr; i++) {
        if (i > 1 && sites[i] == sites[i - 1]) continue; /* note i>1, :3391 */
        buf[nb++] = ((uint64_t)sites[i]) << 35 | i;
    }
    for (uint32_t i = 0; i < nc; i++)
        buf[nb++] = ((uint64_t)(uint32_t)((calls->pos[i] << 3) | 4u | calls->cat[i])) << 32; /* u32 wrap, :3398 */
    qsort(buf, nb, sizeof(uint64_t), cmp_u64);

    const uint64_t callbit = 4ull << 32;
    size_t n_out = 0, out_m = 64;
    uint32_t *res = (uint32_t *)malloc(sizeof(uint32_t) * out_m);
    uint32_t first = UINT32_MAX;
    for (size_t b = 0; b < nb; b++) {
        if (buf[b] & callbit) continue;
        uint32_t pos_i = (uint32_t)buf[b];
        for (uint32_t j = pos_i; j < n_sites; j++) {
            if (sites[j] != sites[pos_i]) break;
            int L = ms->lens[j], n = 0;
            uint32_t key = 0;
            for (size_t t = b; t < nb - 1;) { /* t < buf.n-1, :3420 */
                if (buf[t] & callbit) { t++; continue; }
                if ((buf[t] >> 35) == (buf[t + 1] >> 35) && (buf[t + 1] & callbit)) {
                    uint32_t sym = (uint32_t)(buf[t + 1] >> 32) & 3; /* "mu-"[..] then m=0,u=1,other=2 */
                    key = key << 2 | (sym == 0 ? 0u : sym == 1 ? 1u : 2u);
                    n++; t += 2;
                } else { key = key << 2 | 2u; n++; t++; }
                if (n >= L) break;
            }
            if (n != L) continue;
            if (first == UINT32_MAX) first = j;
            if (n_out == out_m) { out_m *= 2; res = (uint32_t *)realloc(res, sizeof(uint32_t) * out_m); }
            res[n_out++] = key;
        }
    }
    free(buf);
    if (n_out == 0) { free(res); return 0; }
    *out = res;
    *start_i = first;
    return (int)n_out;
}

/* ---------------- count tables ---------------- */

typedef struct { uint32_t *key; uint16_t *c0, *c1; int n, m; uint16_t sum[2]; } site_tab_t;
typedef struct { const port_sites_t *ms; site_tab_t *t; uint32_t min_i, max_i; } tables_t;

static void tab_insert(tables_t *T, const uint32_t *mmr, int n_mmr, uint32_t start_i, int hap) { /* :3453-3486 */
    for (int i0 = 0; i0 < n_mmr; i0++) {
        site_tab_t *s = &T->t[(int)(i0 + start_i)];
        int j;
        for (j = 0; j < s->n; j++) if (s->key[j] == mmr[i0]) break;
        if (j == s->n) {
            if (s->n == s->m) {
                s->m = s->m ? s->m * 2 : 8;
                s->key = (uint32_t *)realloc(s->key, sizeof(uint32_t) * s->m);
                s->c0 = (uint16_t *)realloc(s->c0, sizeof(uint16_t) * s->m);
                s->c1 = (uint16_t *)realloc(s->c1, sizeof(uint16_t) * s->m);
            }
            s->key[j] = mmr[i0]; s->c0[j] = 0; s->c1[j] = 0; s->n++;
        }
        if (hap == 0) s->c0[j]++; else s->c1[j]++;
        s->sum[hap]++;
    }
}

static void tab_update_range(tables_t *T, int cov) { /* :3669-3691 */
    for (int i = (int)T->min_i; i >= 0; i--) {
        int sum = T->t[i].sum[0] + T->t[i].sum[1];
        if (sum >= cov) T->min_i = (uint32_t)i; else break;
    }
    for (int i = (int)T->max_i; i < T->ms->n; i++) {
        if (T->t[i].sum[0] + T->t[i].sum[1] >= cov) T->max_i = (uint32_t)i; else break;
    }
}

/* query_counts_of_mmrs + the summing loop of use_mmr_count_predict_tag_for_one_read, :3487-3515, 3617-3636 */
static void tab_score(const tables_t *T, const uint32_t *mmr, int n_mmr, uint32_t start_i, int hap, float *score,
                      int *score_l) {
    float s = 0;
    int l = 0;
    for (int i0 = 0; i0 < n_mmr; i0++) {
        int i = (int)(start_i + (uint32_t)i0);
        if ((uint32_t)i < T->min_i || (uint32_t)i >= T->max_i) continue; /* int vs uint32 compare => unsigned */
        const site_tab_t *st = &T->t[i];
        for (int j = 0; j < st->n; j++) {
            if (st->key[j] == mmr[i0]) {
                uint32_t cnt = hap == 0 ? st->c0[j] : st->c1[j];
                uint32_t sum = st->sum[hap];
                if (sum != 0) {
                    float v = (float)cnt / sum;
                    l++;
                    if (v > 0) { s += v; l++; }
                }
                break;
            }
        }
    }
    *score = s;
    *score_l = l;
}

static int predict_one(const tables_t *T, const uint32_t *mmr, int n_mmr, uint32_t start_i, float *best) { /* :3594-3656 */
    float s0, s1;
    int l0, l1;
    tab_score(T, mmr, n_mmr, start_i, 0, &s0, &l0);
    tab_score(T, mmr, n_mmr, start_i, 1, &s1, &l1);
    float diff = s0 > s1 ? s0 - s1 : s1 - s0;
    if (diff < 3.0f && (l0 < 3 || l1 < 3)) { *best = 0; return -1; }
    *best = diff;
    return s0 > s1 ? 0 : 1;
}

/* ---------------- greedy propagation ---------------- */

typedef struct { float score; int tag; uint32_t id; } cand_t;

static void stable_sort_cands(cand_t *a, int n) { /* ks_mergesort: stable, ascending by score */
    for (int i = 1; i < n; i++) {
        cand_t x = a[i];
        int j = i - 1;
        while (j >= 0 && x.score < a[j].score) { a[j + 1] = a[j]; j--; }
        a[j + 1] = x;
    }
}

static void greedy(port_readset_t *rs, const port_sites_t *ms, int direction, int n_cand, int cov_run, uint32_t *order,
                   int *n_order) {
    tables_t T;
    T.ms = ms;
    T.t = (site_tab_t *)calloc(ms->n ? ms->n : 1, sizeof(site_tab_t));
    const uint32_t *ref_ids;
    int n_ref;
    if (direction == 0) { /* :3976-3998 */
        T.min_i = 0; T.max_i = 0;
        ref_ids = rs->ids_left; n_ref = (int)rs->n_left;
        for (int i = (int)T.max_i; i < ms->n; i++) { if (ms->real_pos[i] <= rs->ref_start) T.max_i++; else break; }
    } else {
        T.min_i = (uint32_t)(ms->n - 1); T.max_i = (uint32_t)(ms->n - 1);
        ref_ids = rs->ids_right; n_ref = (int)rs->n_right;
        for (int i = (int)T.min_i; i >= 0; i--) { if (ms->real_pos[i] > rs->ref_end) T.min_i--; else break; }
    }
    /* seed, :3776-3810 */
    for (int i = 0; i < n_ref; i++) {
        uint32_t id = ref_ids[i];
        int hap = rs->hp[id];
        if ((hap == 0 || hap == 1) && rs->mmr_start_i[id] != UINT32_MAX)
            tab_insert(&T, rs->mmr[id], rs->mmr_n[id], rs->mmr_start_i[id], hap);
    }
    tab_update_range(&T, cov_run);
    /* untag everything but the ref reads, :4010-4025 — including the (readID<<2)|hp packing quirk */
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * (n_ref ? n_ref : 1));
    for (int i = 0; i < n_ref; i++) tmp[i] = (ref_ids[i] << 2) | (uint32_t)rs->hp[ref_ids[i]];
    for (int i = 0; i < rs->n; i++) rs->hp[i] = 2;
    for (int i = 0; i < n_ref; i++) {
        uint32_t id = tmp[i] >> 2;
        if (id < (uint32_t)rs->n_loaded) rs->hp[id] = (int)(tmp[i] & 3); /* writes past rs->n land in spare slots */
    }
    free(tmp);

    const int n = rs->n;
    int i_last = direction == 0 ? 0 : n - 1;
    const int inc = direction == 0 ? 1 : -1;
    int failed = 0;
    cand_t *cands = (cand_t *)malloc(sizeof(cand_t) * (n_cand > 0 ? n_cand : 1));
    *n_order = 0;
    for (;;) { /* :4032-4071 */
        int nc = 0;
        if ((direction == 0 && i_last >= n) || (direction != 0 && i_last <= 0)) break;
        for (int i0 = i_last; direction == 0 ? i0 < n : i0 >= 0; i0 += inc) {
            int i = direction == 0 ? i0 : (int)(uint32_t)rs->revbuf[i0];
            if (rs->hp[i] != 0 && rs->hp[i] != 1) {
                cands[nc++].id = (uint32_t)i;
                if (nc >= n_cand) break;
            }
        }
        int inserted = 0;
        if (nc > 0) {
            for (int c = 0; c < nc; c++) {
                uint32_t id = cands[c].id;
                cands[c].tag = predict_one(&T, rs->mmr[id], rs->mmr_n[id], rs->mmr_start_i[id], &cands[c].score);
            }
            stable_sort_cands(cands, nc);
            for (int c = nc - 1; c >= 0; c--) {
                uint32_t id = cands[c].id;
                if ((cands[c].tag == 0 || cands[c].tag == 1) && rs->mmr_start_i[id] != UINT32_MAX) {
                    rs->hp[id] = cands[c].tag;
                    tab_insert(&T, rs->mmr[id], rs->mmr_n[id], rs->mmr_start_i[id], cands[c].tag);
                    if (order) order[(*n_order)++] = id;
                    inserted = 1;
                    break;
                }
            }
            if (inserted) tab_update_range(&T, cov_run);
        }
        if (!inserted) {
            failed++;
            if (failed > 10) break;
            i_last += n_cand * inc;
            continue;
        }
        failed = 0;
    }
    free(cands);
    for (int i = 0; i < ms->n; i++) { free(T.t[i].key); free(T.t[i].c0); free(T.t[i].c1); }
    free(T.t);
}

/* ---------------- evaluation ---------------- */

static double lchoose(int n, int k) {
    if (k < 0 || k > n) return -INFINITY;
    return lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n - k + 1.0);
}
double port_fisher_two_sided(int n11, int n12, int n21, int n22) {
    /* kt_fisher_exact (htslib kfunc.c, external): sum of hypergeometric probabilities not larger than
     * the observed table's, with the customary 1e-8 relative slack */
    int r1 = n11 + n12, c1 = n11 + n21, n = n11 + n12 + n21 + n22;
    int hi = c1 < r1 ? c1 : r1, lo = r1 + c1 - n;
    if (lo < 0) lo = 0;
    if (lo == hi) return 1.0;
    double q = exp(lchoose(r1, n11) + lchoose(n - r1, c1 - n11) - lchoose(n, c1)), two = 0;
    for (int x = lo; x <= hi; x++) {
        double p = exp(lchoose(r1, x) + lchoose(n - r1, c1 - x) - lchoose(n, c1));
        if (p < 1.00000001 * q) two += p;
    }
    return two > 1.0 ? 1.0 : two;
}

float port_evaluate_separation(const uint8_t *ref, const uint8_t *query, int n, int *join_dir, int table[4]) {
    int buf[2][2] = {{0, 0}, {0, 0}};
    for (int i = 0; i < n; i++) {
        if (ref[i] != 0 && ref[i] != 1) continue;
        if (query[i] != 0 && query[i] != 1) continue;
        buf[ref[i]][query[i]]++;
    }
    if (table) { table[0] = buf[0][0]; table[1] = buf[0][1]; table[2] = buf[1][0]; table[3] = buf[1][1]; }
#define MIN2(a, b) ((a) <= (b) ? (a) : (b))
    int hard_cov_fail = MIN2(buf[0][0], buf[0][1]) > 15 || MIN2(buf[1][0], buf[1][1]) > 15;
    float scores[2], mn, mx;
    int which_way = 0;
    for (int i = 0; i < 2; i++) {
        if (buf[i][0] > buf[i][1]) { mn = (float)buf[i][1]; mx = (float)buf[i][0]; which_way = i == 0 ? which_way + 1 : which_way - 1; }
        else { mn = (float)buf[i][0]; mx = (float)buf[i][1]; which_way = i == 0 ? which_way - 1 : which_way + 1; }
        if (MIN2(buf[0][0], buf[0][1]) > 5 || MIN2(buf[1][0], buf[1][1]) > 5) { *join_dir = -9; return 1.0f; }
        if (mx == 0) { *join_dir = -9; return 1.0f; }
        mn = mn == 0 ? 1 : mn;
        if (mx / mn < 3) { *join_dir = -9; return 1.0f; }
        scores[i] = mx / mn;
    }
    double p = port_fisher_two_sided(buf[0][0], buf[0][1], buf[1][0], buf[1][1]);
    if (p < 0.001 && !hard_cov_fail) { *join_dir = which_way; return MIN2(scores[0], scores[1]); }
    *join_dir = -9;
    return 1.0f;
}

/* haplotag_region2 with a single permutation, :4088-4214 */
static int region2(port_readset_t *rs, const port_sites_t *ms, int direction, int n_cand, int cov_run, uint8_t *tags_out,
                   uint8_t *prop_out, int table[4], float *score_out, int *which_way_out, uint32_t *order, int *n_order) {
    const int nl = rs->n_loaded;
    uint8_t *initial = (uint8_t *)malloc(nl ? nl : 1);
    for (int i = 0; i < nl; i++) initial[i] = (uint8_t)rs->hp[i];
    greedy(rs, ms, direction, n_cand, cov_run, order, n_order);
    for (int i = 0; i < nl; i++) prop_out[i] = (uint8_t)rs->hp[i]; /* propagated tags, before the summary */
    const uint32_t *sid = direction == 0 ? rs->ids_right_strict : rs->ids_left_strict;
    const int ns = (int)(direction == 0 ? rs->n_right_strict : rs->n_left_strict);
    uint8_t *a = (uint8_t *)malloc(ns ? ns : 1), *b = (uint8_t *)malloc(ns ? ns : 1);
    for (int i = 0; i < ns; i++) { a[i] = initial[sid[i]]; b[i] = (uint8_t)rs->hp[sid[i]]; }
    int which_way = 0;
    float score = port_evaluate_separation(a, b, ns, &which_way, table);
    free(a); free(b);
    *score_out = score;
    *which_way_out = which_way;
    int ret = -1;
    if (score >= 2 && which_way != 0) ret = which_way > 0 ? 0 : 1;
    /* noperm branch: success keeps the propagated tags, failure => all unphased (:4189-4205) */
    if (ret < 0) for (int i = 0; i < rs->n; i++) rs->hp[i] = 2;
    for (int i = 0; i < nl; i++) tags_out[i] = (uint8_t)rs->hp[i];
    free(initial);
    return ret;
}

/* ---------------- window driver ---------------- */

static void store_mmrs(port_readset_t *rs, const port_sites_t *ms) { /* :3518-3550 */
    for (int i = 0; i < rs->n; i++) {
        uint32_t *m = NULL, st = 0;
        int n = port_mmr_of_read(&rs->calls[i], ms, &m, &st);
        if (n == 0 || st == UINT32_MAX) { rs->mmr[i] = NULL; rs->mmr_n[i] = 0; rs->mmr_start_i[i] = 0; free(m); }
        else { rs->mmr[i] = m; rs->mmr_n[i] = n; rs->mmr_start_i[i] = st; }
    }
}

port_window_t *port_window_run(const pomfret_gpu_read_desc *reads, int n_in, uint32_t ref_start, uint32_t ref_end,
                               const pomfret_gpu_config *cfg) {
    port_window_t *w = (port_window_t *)calloc(1, sizeof(port_window_t));
    port_readset_t *rs = (port_readset_t *)calloc(1, sizeof(port_readset_t));
    w->rs = rs;
    w->decision = w->join_fwd = w->join_bwd = -1;
    w->read_ids = (int32_t *)malloc(sizeof(int32_t) * (n_in ? n_in : 1));
    w->status = (uint32_t *)calloc(n_in ? n_in : 1, sizeof(uint32_t));
    rs->ref_start = ref_start; /* itvl_s >= 0 always: the seam passes uint32 */
    rs->ref_end = ref_end;
    int cap = n_in ? n_in : 1;
    rs->hp = (int *)malloc(sizeof(int) * cap);
    rs->strand = (uint8_t *)malloc(cap);
    rs->start_pos = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    rs->end_pos = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    rs->calls = (port_calls_t *)calloc(cap, sizeof(port_calls_t));
    rs->revbuf = (uint64_t *)malloc(sizeof(uint64_t) * cap);
    rs->ids_left = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    rs->ids_left_strict = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    rs->ids_right = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    rs->ids_right_strict = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    rs->mmr = (uint32_t **)calloc(cap, sizeof(uint32_t *));
    rs->mmr_n = (int *)calloc(cap, sizeof(int));
    rs->mmr_start_i = (uint32_t *)calloc(cap, sizeof(uint32_t));

    /* load_reads_given_interval after the record filters, :1087-1163 */
    int n = 0, cov_check[2] = {0, 0};
    const int itvl_s = (int)ref_start, itvl_e = (int)ref_end;
    for (int i = 0; i < n_in; i++) {
        uint32_t st = 0, endp = 0;
        port_calls_t c = {0, 0, 0, 0};
        int rc = port_decode_read(&reads[i], cfg->lo, cfg->hi, &c, &st, &endp);
        w->status[i] = st;
        w->read_ids[i] = -1;
        if (rc == POMFRET_GPU_ERR_FATAL_CIGAR) { w->status_code = rc; free(c.pos); free(c.cat); continue; }
        if (rc != 1) { free(c.pos); free(c.cat); continue; }
        w->read_ids[i] = n;
        rs->calls[n] = c;
        rs->hp[n] = reads[i].hp;
        rs->strand[n] = (reads[i].flag & 16) != 0;
        rs->start_pos[n] = reads[i].pos;
        rs->end_pos[n] = endp;
        uint32_t start_pos = reads[i].pos;
        uint64_t end_pos = endp;
        rs->revbuf[n] = (end_pos << 32) | (uint32_t)n;
        /* uint32 start_pos vs int itvl_s: usual arithmetic conversions => unsigned compare, :1127 */
        if (start_pos <= (uint32_t)itvl_s) {
            rs->ids_left[rs->n_left++] = (uint32_t)n;
            if (end_pos > (uint64_t)(int64_t)itvl_s) rs->ids_left_strict[rs->n_left_strict++] = (uint32_t)n;
            if (rs->hp[n] == 0 || rs->hp[n] == 1) cov_check[rs->hp[n]]++;
        } else if (end_pos >= (uint64_t)(int64_t)itvl_e) {
            rs->ids_right[rs->n_right++] = (uint32_t)n;
            if (start_pos < (uint32_t)itvl_e) rs->ids_right_strict[rs->n_right_strict++] = (uint32_t)n;
        }
        n++;
    }
    qsort(rs->revbuf, n, sizeof(uint64_t), cmp_u64);
    rs->n_loaded = n;
    rs->n = (cov_check[0] < 15 || cov_check[1] < 15) ? 0 : n; /* :1161-1163 */
    w->n_reads = rs->n;
    w->n_reads_loaded = n;
    w->tags_final = (uint8_t *)malloc(n ? n : 1);
    w->tags_fwd = (uint8_t *)malloc(n ? n : 1);
    w->tags_bwd = (uint8_t *)malloc(n ? n : 1);
    w->prop_fwd = (uint8_t *)malloc(n ? n : 1);
    w->prop_bwd = (uint8_t *)malloc(n ? n : 1);
    w->order_fwd = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    w->order_bwd = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    w->mmr_bwd = (uint32_t **)calloc(cap, sizeof(uint32_t *));
    w->mmr_n_bwd = (int *)calloc(cap, sizeof(int));
    w->mmr_start_bwd = (uint32_t *)calloc(cap, sizeof(uint32_t));

    port_sites(rs, cfg->cov_for_selection, cfg->k, cfg->k_span, 0, &w->sites[0]);
    port_sites(rs, cfg->cov_for_selection, cfg->k, cfg->k_span, 1, &w->sites[1]);
    w->n_sites_fwd = w->sites[0].n;
    w->n_sites_bwd = w->sites[1].n;
    for (int i = 0; i < n; i++) w->tags_final[i] = w->tags_fwd[i] = w->tags_bwd[i] = w->prop_fwd[i] = w->prop_bwd[i] = (uint8_t)rs->hp[i];
    if (w->sites[0].n == 0 || w->sites[1].n == 0) return w; /* :4266-4270 */

    uint8_t *initial = (uint8_t *)malloc(n ? n : 1);
    for (int i = 0; i < n; i++) initial[i] = (uint8_t)rs->hp[i];
    store_mmrs(rs, &w->sites[1]);
    for (int i = 0; i < rs->n; i++) { w->mmr_bwd[i] = rs->mmr[i]; w->mmr_n_bwd[i] = rs->mmr_n[i]; w->mmr_start_bwd[i] = rs->mmr_start_i[i]; }
    w->join_bwd = region2(rs, &w->sites[1], 1, cfg->n_candidates_per_iter, cfg->cov_for_runtime, w->tags_bwd,
                          w->prop_bwd, w->table_bwd, &w->score_bwd, &w->which_way_bwd, w->order_bwd, &w->n_order_bwd);
    for (int i = 0; i < n; i++) rs->hp[i] = initial[i]; /* do_reset=1 */
    store_mmrs(rs, &w->sites[0]);
    w->join_fwd = region2(rs, &w->sites[0], 0, cfg->n_candidates_per_iter, cfg->cov_for_runtime, w->tags_fwd,
                          w->prop_fwd, w->table_fwd, &w->score_fwd, &w->which_way_fwd, w->order_fwd, &w->n_order_fwd);
    if (w->join_fwd != w->join_bwd || (w->join_fwd == -1 && w->join_bwd == -1)) {
        for (int i = 0; i < rs->n; i++) rs->hp[i] = 2;
        w->decision = -1;
    } else w->decision = w->join_fwd;
    for (int i = 0; i < n; i++) w->tags_final[i] = (uint8_t)rs->hp[i];
    free(initial);
    return w;
}

void port_window_free(port_window_t *w) {
    if (!w) return;
    port_readset_t *rs = w->rs;
    for (int i = 0; i < rs->n_loaded; i++) {
        free(rs->calls[i].pos); free(rs->calls[i].cat);
        free(rs->mmr[i]);
        free(w->mmr_bwd[i]);
    }
    free(rs->hp); free(rs->strand); free(rs->start_pos); free(rs->end_pos); free(rs->calls); free(rs->revbuf);
    free(rs->ids_left); free(rs->ids_left_strict); free(rs->ids_right); free(rs->ids_right_strict);
    free(rs->mmr); free(rs->mmr_n); free(rs->mmr_start_i); free(rs);
    port_sites_free(&w->sites[0]); port_sites_free(&w->sites[1]);
    free(w->tags_final); free(w->tags_fwd); free(w->tags_bwd); free(w->prop_fwd); free(w->prop_bwd); free(w->read_ids); free(w->status);
    free(w->mmr_bwd); free(w->mmr_n_bwd); free(w->mmr_start_bwd); free(w->order_fwd); free(w->order_bwd);
    free(w);
}
