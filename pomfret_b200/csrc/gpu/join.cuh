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

constexpr int JOIN_THREADS = 512;
constexpr int JOIN_WARPS = JOIN_THREADS / 32;
constexpr int JOIN_MAX_CAND = 128;
constexpr int JOIN_U = 8;                 // methmer sub-chunks (32 each) whose loads are issued together
constexpr int JOIN_CHUNK = JOIN_U * 32;   // values staged per warp before the ordered sum

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

// Range growth of update_available_methmer_range (blockjoin.c:3669-3691), warp-parallel: the left edge
// walks down from `mn` while the site's coverage reaches cov, the right edge walks up from `mx`.
__device__ __forceinline__ void grow_range(uint32_t *tab, uint32_t tab_base, uint32_t row_words, uint32_t n_keys,
                                           uint32_t n_sites, int cov, uint32_t &mn, uint32_t &mx) {
    const int lane = (int)lane_id();
    for (int i0 = (int)mn;; i0 -= 32) {
        const int i = i0 - lane;
        bool ok = false;
        if (i >= 0) {
            const uint32_t sm = site_row(tab, tab_base, (uint32_t)i, row_words)[n_keys];
            ok = (int)((sm & 0xffffu) + (sm >> 16)) >= cov;
        }
        const unsigned bad = ~__ballot_sync(FULL_MASK, ok);  // first lane that stops the walk (or ran past site 0)
        const int f = bad ? __ffs((int)bad) - 1 : 32;
        if (f > 0) mn = (uint32_t)(i0 - (f - 1));
        if (f < 32) break;
    }
    for (int i0 = (int)mx;; i0 += 32) {
        const int i = i0 + lane;
        bool ok = false;
        if (i < (int)n_sites) {
            const uint32_t sm = site_row(tab, tab_base, (uint32_t)i, row_words)[n_keys];
            ok = (int)((sm & 0xffffu) + (sm >> 16)) >= cov;
        }
        const unsigned bad = ~__ballot_sync(FULL_MASK, ok);
        const int f = bad ? __ffs((int)bad) - 1 : 32;
        if (f > 0) mx = (uint32_t)(i0 + (f - 1));
        if (f < 32) break;
    }
}

__global__ void __launch_bounds__(JOIN_THREADS) join_kernel(JoinParams P) {
    __shared__ int s_i_last, s_failed, s_done, s_ncand, s_cursor;
    __shared__ uint32_t s_min, s_max;
    __shared__ uint32_t s_cand[JOIN_MAX_CAND];
    __shared__ float s_score[JOIN_MAX_CAND];
    __shared__ int s_tag[JOIN_MAX_CAND];
    __shared__ float s_val[JOIN_WARPS][2][JOIN_CHUNK];
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
    const uint32_t *rev = P.rs_rev + first;
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
    // (every read touches each site at most once, and the u16 halves add independently: order is irrelevant)
    for (uint32_t r = warp; r < n_ref; r += JOIN_WARPS) {
        const uint32_t id = ref_ids[r];
        const int hap = P.rs_hp[first + id];
        if (hap == 0 || hap == 1) {
            const uint32_t nm = mm_n[id], st = mm_start[id], off = mm_off[id];
            const uint32_t inc = hap == 0 ? 1u : 0x10000u;
            for (uint32_t i0 = lane; i0 < nm; i0 += 32) {
                uint32_t *row = site_row(tab, tab_base, st + i0, row_words);
                atomicAdd(&row[pool[off + i0]], inc);
                atomicAdd(&row[n_keys], inc);
            }
        }
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
    }
    __syncthreads();

    // ---- extension loop, :4032-4071 ----
    // Warp 0 owns the loop state between the barriers: available range, candidate list, failure count.
    // The candidate list ("the first n_cand untagged reads in scan order from i_last", :4039-4045) is kept
    // incrementally: a success removes the tagged read and appends the next untagged one behind the scan
    // cursor; a failure moves i_last (:4064-4068) and rebuilds it.
    uint32_t n_order = 0;
    int nc = 0, cursor = 0;          // warp 0 only (uniform)
    bool rebuild = true, grow = true;  // first pass: update_available_methmer_range after seeding, fresh list
    int last_best = -1;
    for (;;) {
        if (warp == 0) {
            uint32_t mn = s_min, mx = s_max;
            if (grow) { grow_range(tab, tab_base, row_words, n_keys, n_sites, P.cov_run, mn, mx); if (lane == 0) { s_min = mn; s_max = mx; } }
            const int i_last = s_i_last;
            const bool done = (d == 0 && i_last >= (int)n) || (d != 0 && i_last <= 0);
            if (!done) {
                if (rebuild) { nc = 0; cursor = i_last; }
                else if (last_best >= 0) {
                    // drop entry last_best, keep the order of the rest
                    for (int c0 = 0; c0 < nc; c0 += 32) {
                        const int c = c0 + (int)lane;
                        uint32_t v = 0;
                        const bool mv = c > last_best && c < nc;
                        if (mv) v = s_cand[c];
                        __syncwarp();
                        if (mv) s_cand[c - 1] = v;
                    }
                    nc--;
                    __syncwarp();
                }
                // refill from the cursor
                while (nc < n_cand) {
                    const int i0 = d == 0 ? cursor + (int)lane : cursor - (int)lane;
                    const bool in = d == 0 ? i0 < (int)n : i0 >= 0;
                    uint32_t id = 0;
                    bool unt = false;
                    if (in) {
                        id = d == 0 ? (uint32_t)i0 : rev[i0];
                        const uint8_t t = tags[id];
                        unt = t != 0 && t != 1;
                    }
                    const unsigned um = __ballot_sync(FULL_MASK, unt);
                    const int rank = __popc(um & ((1u << lane) - 1u));
                    const int room = n_cand - nc;
                    if (unt && rank < room) s_cand[nc + rank] = id;
                    const int found = __popc(um);
                    if (found >= room) {
                        // the list is full: the cursor stops right behind the read that filled it
                        const int fill_lane = (int)__fns(um, 0, room);
                        cursor += d == 0 ? fill_lane + 1 : -(fill_lane + 1);
                        nc = n_cand;
                        break;
                    }
                    nc += found;
                    cursor += d == 0 ? 32 : -32;
                    if (__ballot_sync(FULL_MASK, in) != FULL_MASK) break;  // ran past the last read
                }
                __syncwarp();
            }
            if (lane == 0) { s_ncand = nc; if (done) s_done = 1; }
        }
        __syncthreads();
        if (s_done) break;
        const int ncand = s_ncand;
        const uint32_t rmin = s_min, rmax = s_max;
        // ---- score the candidates, one warp each (use_mmr_count_predict_tag_for_one_read, :3594-3656) ----
        for (int c = (int)warp; c < ncand; c += JOIN_WARPS) {
            const uint32_t id = s_cand[c];
            const uint32_t nm = mm_n[id], st = mm_start[id], off = mm_off[id];
            float sc_h = 0.f;  // lane 0: haplotype 0, lane 1: haplotype 1
            int l0 = 0, l1 = 0;
            for (uint32_t base = 0; base < nm; base += JOIN_CHUNK) {
                uint32_t key[JOIN_U], cnt[JOIN_U], sums[JOIN_U];
                bool inr[JOIN_U];
#pragma unroll
                for (int u = 0; u < JOIN_U; u++) {
                    const uint32_t i0 = base + u * 32 + lane;
                    const uint32_t site = st + i0;
                    inr[u] = i0 < nm && !(site < rmin || site >= rmax);
                    key[u] = inr[u] ? pool[off + i0] : 0u;
                }
#pragma unroll
                for (int u = 0; u < JOIN_U; u++) {
                    cnt[u] = 0; sums[u] = 0;
                    if (inr[u]) {
                        const uint32_t *row = site_row(tab, tab_base, st + base + u * 32 + lane, row_words);
                        cnt[u] = row[key[u]];
                        sums[u] = row[n_keys];
                    }
                }
                int n0 = 0, n1 = 0;  // non-zero terms staged so far (adding +0.0f is exact, so zeros are skipped)
#pragma unroll
                for (int u = 0; u < JOIN_U; u++) {
                    if (base + u * 32 >= nm) break;
                    float v0 = 0.f, v1 = 0.f;
                    bool p0 = false, p1 = false;
                    if (cnt[u] != 0) {  // key present at this site
                        const uint32_t sum0 = sums[u] & 0xffffu, sum1 = sums[u] >> 16;
                        if (sum0 != 0) { p0 = true; v0 = __fdiv_rn((float)(cnt[u] & 0xffffu), (float)sum0); }
                        if (sum1 != 0) { p1 = true; v1 = __fdiv_rn((float)(cnt[u] >> 16), (float)sum1); }
                    }
                    const unsigned z0 = __ballot_sync(FULL_MASK, v0 > 0.f), z1 = __ballot_sync(FULL_MASK, v1 > 0.f);
                    l0 += __popc(__ballot_sync(FULL_MASK, p0)) + __popc(z0);
                    l1 += __popc(__ballot_sync(FULL_MASK, p1)) + __popc(z1);
                    const unsigned lt = (1u << lane) - 1u;
                    if (v0 > 0.f) s_val[warp][0][n0 + __popc(z0 & lt)] = v0;
                    if (v1 > 0.f) s_val[warp][1][n1 + __popc(z1 & lt)] = v1;
                    n0 += __popc(z0);
                    n1 += __popc(z1);
                }
                __syncwarp();
                if (lane < 2) {  // strictly in methmer order, IEEE round-to-nearest (blockjoin.c:3620-3636)
                    const int nn = lane == 0 ? n0 : n1;
                    const float *v = s_val[warp][lane];
#pragma unroll 4
                    for (int t = 0; t < nn; t++) sc_h = __fadd_rn(sc_h, v[t]);
                }
                __syncwarp();
            }
            const float score1 = __shfl_sync(FULL_MASK, sc_h, 1);
            const float score0 = __shfl_sync(FULL_MASK, sc_h, 0);
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
        // ---- stable ascending sort + scan from the top == max score, ties to the later candidate (:3729-3760);
        //      every warp finds it on its own ----
        int best = -1;
        {
            unsigned long long bk = 0;
            for (int c = (int)lane; c < ncand; c += 32) {
                const int t = s_tag[c];
                if (t == 0 || t == 1) {
                    const unsigned long long k = (((unsigned long long)__float_as_uint(s_score[c]) << 32) | (unsigned)c) + 1ull;
                    bk = k > bk ? k : bk;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long t = __shfl_xor_sync(FULL_MASK, bk, o);
                bk = t > bk ? t : bk;
            }
            if (bk) best = (int)((bk - 1ull) & 0xffffffffull);
        }
        if (best >= 0) {
            const uint32_t id = s_cand[best];
            const int hap = s_tag[best];
            const uint32_t nm = mm_n[id], st = mm_start[id], off = mm_off[id];
            const uint32_t inc = hap == 0 ? 1u : 0x10000u;
            for (uint32_t i0 = tid; i0 < nm; i0 += JOIN_THREADS) {
                uint32_t *row = site_row(tab, tab_base, st + i0, row_words);
                row[pool[off + i0]] += inc;
                row[n_keys] += inc;
            }
            if (tid == 0) {
                tags[id] = (uint8_t)hap;
                P.order[d][first + n_order] = id;
            }
            n_order++;
        }
        if (warp == 0) {
            last_best = best;
            if (best >= 0) { rebuild = false; grow = true; if (lane == 0) s_failed = 0; }
            else {
                rebuild = true; grow = false;
                if (lane == 0) {
                    s_failed++;
                    if (s_failed > 10) s_done = 1;
                    s_i_last += d == 0 ? n_cand : -n_cand;
                }
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
