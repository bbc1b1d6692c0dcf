// (e) Greedy haplotag propagation across a phase-block gap, one CTA per (window, direction).
//
// Replaces haplotag_region1 (reference blockjoin.c:3958-4080) with everything it calls:
//   insert_ref_reads_methmer_counts 3776-3810, insert_mmrs_to_counts 3453-3486,
//   query_counts_of_mmrs 3487-3515, use_mmr_count_predict_tag_for_one_read 3594-3656,
//   update_available_methmer_range 3669-3691, predict_tags_of_reads 3693-3774,
// and tabulates the 2x2 table of evaluate_separation1 (3881-3893).  Fisher's test and the join rule
// stay on the host (fp64 lgamma, a handful of integers).
//
// The reference keeps, per site, a growing list of (key, count per haplotype); membership is all that
// the list adds over a dense table, and "key present" == "some haplotype count is non-zero", so the
// table here is dense: one 32-bit word per (site, key) holding both 16-bit counts, one word per site
// for the two sums.  It lives in global memory (L1/L2 resident, a few hundred KB per window).
//
// Score sums are order sensitive fp32 (blockjoin.c:3620-3636): values are produced in parallel, one
// lane per methmer, then added strictly in methmer order by one lane per haplotype with IEEE
// round-to-nearest adds and divides (no fast-math, no FMA contraction possible).
#ifndef POMFRET_GPU_JOIN_CUH
#define POMFRET_GPU_JOIN_CUH
#include "gpu_rt.h"
#include "types.h"

namespace pomfret_gpu {

constexpr int JOIN_THREADS = 256;
constexpr int JOIN_WARPS = JOIN_THREADS / 32;
constexpr int JOIN_MAX_CAND = 128;

struct JoinParams {
    const WindowRec *win;
    WindowState *state;
    const uint32_t *rs_rev;
    const int32_t *rs_hp;  // initial tags per slot
    const uint32_t *ids_left, *ids_left_strict, *ids_right, *ids_right_strict;
    const uint32_t *site_pos;
    const uint32_t *mm_off[2], *mm_n[2], *mm_start[2];
    const uint32_t *mmr_pool;
    uint32_t *tab;         // table pool: per site (n_keys + 1) words
    uint8_t *tags[2];      // propagated tags per slot and direction
    uint32_t *order[2];    // tagging order per direction (slot-indexed slices)
    int32_t n_cand;
    int32_t cov_run;
    int32_t k;
};

__device__ __forceinline__ uint32_t *site_row(uint32_t *tab, uint32_t tab_base, uint32_t site, uint32_t row_words) {
    return tab + ((size_t)tab_base + site) * row_words;
}

__global__ void __launch_bounds__(JOIN_THREADS) join_kernel(JoinParams P) {
    __shared__ int s_i_last, s_failed, s_done, s_ncand, s_best, s_best_tag;
    __shared__ uint32_t s_min, s_max;
    __shared__ uint32_t s_cand[JOIN_MAX_CAND];
    __shared__ float s_score[JOIN_MAX_CAND];
    __shared__ int s_tag[JOIN_MAX_CAND];
    __shared__ float s_val[JOIN_WARPS][2][32];
    __shared__ int s_tbl[4];

    const uint32_t w = blockIdx.x >> 1, d = blockIdx.x & 1u;
    const WindowRec W = P.win[w];
    WindowState &S = P.state[w];
    const uint32_t n = S.n, n_sites = S.n_sites;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (n == 0 || n_sites == 0 || S.status != 0) return;
    const uint32_t first = W.first_read;
    const uint32_t n_keys = 1u << (2 * P.k);
    const uint32_t row_words = n_keys + 1;
    const uint32_t tab_base = S.tab_base[d];
    uint32_t *tab = P.tab;
    uint8_t *tags = P.tags[d] + first;
    const uint32_t *mm_off = P.mm_off[d] + first, *mm_n = P.mm_n[d] + first, *mm_start = P.mm_start[d] + first;
    const uint32_t *pool = P.mmr_pool;
    const uint32_t *site_pos = P.site_pos + W.site_off;
    const uint32_t *ref_ids = (d == 0 ? P.ids_left : P.ids_right) + first;
    const uint32_t n_ref = d == 0 ? S.n_left : S.n_right;
    const int n_cand = P.n_cand;

    // ---- wipe the tables (insert_ref_reads_methmer_counts, :3780-3789) ----
    {
        uint32_t *t0 = site_row(tab, tab_base, 0, row_words);
        const size_t words = (size_t)n_sites * row_words;
        for (size_t i = tid; i < words; i += JOIN_THREADS) t0[i] = 0;
    }
    // ---- available range, :3976-4004 ----
    if (tid == 0) {
        uint32_t mn, mx;
        if (d == 0) {
            mn = 0; mx = 0;
            for (int i = 0; i < (int)n_sites; i++) { if (site_pos[i] <= W.ref_start) mx++; else break; }
        } else {
            mn = n_sites - 1; mx = n_sites - 1;
            for (int i = (int)mn; i >= 0; i--) { if (site_pos[i] > W.ref_end) mn--; else break; }
        }
        s_min = mn; s_max = mx;
        s_failed = 0; s_done = 0;
        s_i_last = d == 0 ? 0 : (int)n - 1;
        s_tbl[0] = s_tbl[1] = s_tbl[2] = s_tbl[3] = 0;
    }
    __syncthreads();
    // ---- seed with the reference reads of the starting side, in list order, :3793-3803 ----
    for (uint32_t r = 0; r < n_ref; r++) {
        const uint32_t id = ref_ids[r];
        const int hap = P.rs_hp[first + id];
        if (hap == 0 || hap == 1) {
            const uint32_t nm = mm_n[id], st = mm_start[id], off = mm_off[id];
            for (uint32_t i0 = tid; i0 < nm; i0 += JOIN_THREADS) {
                uint32_t *row = site_row(tab, tab_base, st + i0, row_words);
                row[pool[off + i0]] += hap == 0 ? 1u : 0x10000u;
                row[n_keys] += hap == 0 ? 1u : 0x10000u;
            }
        }
        __syncthreads();
    }
    // ---- un-tag everything but the reference reads, :4010-4025 (with the (id<<2)|hp packing) ----
    for (uint32_t i = tid; i < n; i += JOIN_THREADS) tags[i] = 2;
    __syncthreads();
    if (tid == 0) {
        for (uint32_t r = 0; r < n_ref; r++) {
            uint32_t id = ref_ids[r];
            uint32_t packed = (id << 2) | (uint32_t)P.rs_hp[first + id];
            uint32_t tid2 = packed >> 2;
            if (tid2 < n) tags[tid2] = (uint8_t)(packed & 3u);
        }
        // update_available_methmer_range(cov_for_runtime), :3669-3691
        uint32_t mn = s_min, mx = s_max;
        for (int i = (int)mn; i >= 0; i--) {
            uint32_t sm = site_row(tab, tab_base, (uint32_t)i, row_words)[n_keys];
            if ((int)((sm & 0xffffu) + (sm >> 16)) >= P.cov_run) mn = (uint32_t)i; else break;
        }
        for (int i = (int)mx; i < (int)n_sites; i++) {
            uint32_t sm = site_row(tab, tab_base, (uint32_t)i, row_words)[n_keys];
            if ((int)((sm & 0xffffu) + (sm >> 16)) >= P.cov_run) mx = (uint32_t)i; else break;
        }
        s_min = mn; s_max = mx;
    }
    __syncthreads();

    // ---- extension loop, :4032-4071 ----
    uint32_t n_order = 0;
    for (;;) {
        // 1) candidates: the first n_cand untagged reads in scan order from i_last (warp 0)
        if (warp == 0) {
            int i_last = s_i_last;
            int nc = 0;
            bool done = (d == 0 && i_last >= (int)n) || (d != 0 && i_last <= 0);
            if (!done) {
                int pos = i_last;
                while (nc < n_cand) {
                    int i0 = d == 0 ? pos + (int)lane : pos - (int)lane;
                    bool in = d == 0 ? i0 < (int)n : i0 >= 0;
                    uint32_t id = 0;
                    bool unt = false;
                    if (in) {
                        id = d == 0 ? (uint32_t)i0 : P.rs_rev[first + i0];
                        uint8_t t = tags[id];
                        unt = t != 0 && t != 1;
                    }
                    unsigned um = __ballot_sync(FULL_MASK, unt);
                    int rank = __popc(um & ((1u << lane) - 1u));
                    if (unt && nc + rank < n_cand) s_cand[nc + rank] = id;
                    nc += __popc(um);
                    if (nc > n_cand) nc = n_cand;
                    unsigned inm = __ballot_sync(FULL_MASK, in);
                    if (inm != FULL_MASK) break;
                    pos += d == 0 ? 32 : -32;
                }
            }
            if (lane == 0) { s_ncand = nc; s_done = done ? 1 : 0; }
        }
        __syncthreads();
        if (s_done) break;
        const int ncand = s_ncand;
        const uint32_t rmin = s_min, rmax = s_max;
        // 2) score the candidates, one warp each
        for (int c = (int)warp; c < ncand; c += JOIN_WARPS) {
            const uint32_t id = s_cand[c];
            const uint32_t nm = mm_n[id], st = mm_start[id], off = mm_off[id];
            float score0 = 0.f, score1 = 0.f;  // live in lane 0 / lane 1
            int l0 = 0, l1 = 0;
            for (uint32_t base = 0; base < nm; base += 32) {
                const uint32_t i0 = base + lane;
                float v0 = 0.f, v1 = 0.f;
                bool p0 = false, p1 = false;
                if (i0 < nm) {
                    const uint32_t site = st + i0;
                    if (!(site < rmin || site >= rmax)) {
                        const uint32_t *row = site_row(tab, tab_base, site, row_words);
                        const uint32_t cnt = row[pool[off + i0]];
                        if (cnt != 0) {  // key present at this site
                            const uint32_t sums = row[n_keys];
                            const uint32_t sum0 = sums & 0xffffu, sum1 = sums >> 16;
                            if (sum0 != 0) { p0 = true; v0 = __fdiv_rn((float)(cnt & 0xffffu), (float)sum0); }
                            if (sum1 != 0) { p1 = true; v1 = __fdiv_rn((float)(cnt >> 16), (float)sum1); }
                        }
                    }
                }
                l0 += __popc(__ballot_sync(FULL_MASK, p0)) + __popc(__ballot_sync(FULL_MASK, v0 > 0.f));
                l1 += __popc(__ballot_sync(FULL_MASK, p1)) + __popc(__ballot_sync(FULL_MASK, v1 > 0.f));
                s_val[warp][0][lane] = v0;
                s_val[warp][1][lane] = v1;
                __syncwarp();
                if (lane < 2) {
                    float s = lane == 0 ? score0 : score1;
                    const float *v = s_val[warp][lane];
#pragma unroll 8
                    for (int t = 0; t < 32; t++) s = __fadd_rn(s, v[t]);  // + 0.0f is exact for skipped entries
                    if (lane == 0) score0 = s; else score1 = s;
                }
                __syncwarp();
            }
            score1 = __shfl_sync(FULL_MASK, score1, 1);
            score0 = __shfl_sync(FULL_MASK, score0, 0);
            if (lane == 0) {
                float diff = score0 > score1 ? __fsub_rn(score0, score1) : __fsub_rn(score1, score0);
                int tag;
                float sc;
                if (diff < 3.0f && (l0 < 3 || l1 < 3)) { tag = -1; sc = 0.f; }
                else { tag = score0 > score1 ? 0 : 1; sc = diff; }
                s_score[c] = sc;
                s_tag[c] = tag;
            }
        }
        __syncthreads();
        // 3) stable ascending sort + scan from the top == max score, ties to the later candidate
        if (tid == 0) {
            int best = -1;
            float bs = 0.f;
            for (int c = 0; c < ncand; c++) {
                if (s_tag[c] == 0 || s_tag[c] == 1) {
                    if (best < 0 || s_score[c] >= bs) { best = c; bs = s_score[c]; }
                }
            }
            s_best = best;
            s_best_tag = best >= 0 ? s_tag[best] : -1;
        }
        __syncthreads();
        const int best = s_best;
        if (best >= 0) {
            const uint32_t id = s_cand[best];
            const int hap = s_best_tag;
            const uint32_t nm = mm_n[id], st = mm_start[id], off = mm_off[id];
            for (uint32_t i0 = tid; i0 < nm; i0 += JOIN_THREADS) {
                uint32_t *row = site_row(tab, tab_base, st + i0, row_words);
                row[pool[off + i0]] += hap == 0 ? 1u : 0x10000u;
                row[n_keys] += hap == 0 ? 1u : 0x10000u;
            }
            if (tid == 0) {
                tags[id] = (uint8_t)hap;
                P.order[d][first + n_order] = id;
            }
            n_order++;
        }
        __syncthreads();
        if (tid == 0) {
            if (best >= 0) {
                uint32_t mn = s_min, mx = s_max;
                for (int i = (int)mn; i >= 0; i--) {
                    uint32_t sm = site_row(tab, tab_base, (uint32_t)i, row_words)[n_keys];
                    if ((int)((sm & 0xffffu) + (sm >> 16)) >= P.cov_run) mn = (uint32_t)i; else break;
                }
                for (int i = (int)mx; i < (int)n_sites; i++) {
                    uint32_t sm = site_row(tab, tab_base, (uint32_t)i, row_words)[n_keys];
                    if ((int)((sm & 0xffffu) + (sm >> 16)) >= P.cov_run) mx = (uint32_t)i; else break;
                }
                s_min = mn; s_max = mx;
                s_failed = 0;
            } else {
                s_failed++;
                if (s_failed > 10) s_done = 1;
                s_i_last += d == 0 ? n_cand : -n_cand;
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    // ---- 2x2 table over the far-side strict reads, :3888-3893 and :3940-3951 ----
    {
        const uint32_t *sid = (d == 0 ? P.ids_right_strict : P.ids_left_strict) + first;
        const uint32_t ns = d == 0 ? S.n_right_strict : S.n_left_strict;
        for (uint32_t i = tid; i < ns; i += JOIN_THREADS) {
            const uint32_t id = sid[i];
            const uint8_t ref = (uint8_t)P.rs_hp[first + id];
            const uint8_t q = tags[id];
            if ((ref == 0 || ref == 1) && (q == 0 || q == 1)) atomicAdd(&s_tbl[ref * 2 + q], 1);
        }
        __syncthreads();
        if (tid == 0) {
            for (int t = 0; t < 4; t++) S.table[d][t] = s_tbl[t];
            S.n_order[d] = n_order;
        }
    }
}

}  // namespace pomfret_gpu
#endif
