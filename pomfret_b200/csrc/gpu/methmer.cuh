// (d) Per-read methmer extraction for both directions.
//
// Replaces get_mmr_of_read / store_mmr_of_reads (reference blockjoin.c:3357-3451, 3518-3550) and the
// searches of blockjoin.c:339-421.  The reference merges, per read, the window's methmer start
// positions with the read's calls through a packed 64-bit radix sort and then walks the merged buffer.
// Here the merge is implicit: site starts are already sorted, so each site entry looks its symbol up in
// the read's sorted call list (binary search), and the walk becomes "symbols of the next L entries".
// The reference's quirks are kept: searches run on sites_starts, the equal site is excluded on the right,
// duplicate starts are dropped only for index > 1, and every methmer sharing a start is emitted.
//
// methmer_size_kernel  one CTA per window: [x_left, x_right) per (read, direction), an upper bound of the
//                      methmer count, window-local offsets by a block scan, pool bases by one atomicAdd.
// methmer_fill_kernel  one warp per (read, direction): entries + symbols, then keys.
#ifndef POMFRET_GPU_METHMER_CUH
#define POMFRET_GPU_METHMER_CUH
#include "gpu_rt.h"
#include "types.h"
#include "readset.cuh"

namespace pomfret_gpu {

struct MethmerParams {
    const WindowRec *win;
    WindowState *state;
    const uint32_t *read_win;      // per slot (= batch read index space): window
    const ReadRec *reads;
    const uint32_t *rs_src;
    const uint32_t *r_ncalls, *r_status;
    const uint32_t *calls_pos;
    const uint8_t *calls_cat;
    const uint32_t *site_start[2];
    const uint8_t *site_len[2];
    uint32_t *mm_xl[2], *mm_xr[2], *mm_off[2], *mm_n[2], *mm_start[2];
    uint32_t *pool_total;          // [0] methmer slots, [1] table sites, [2] largest site count of a window, [3] fill queue head
    const uint32_t *order;         // fill queue order: slots by descending record length (nullptr: slot order)
    uint32_t *mmr_pool;            // keys
    uint32_t *ent_pool;            // scratch: (site index << 2 | symbol) per entry
    uint32_t pool_cap;
    uint32_t n_slots;
    int32_t k;
};

// search_arr1 + search_arr(which_end = 0), blockjoin.c:339-421
__device__ inline int search_left(const uint32_t *a, uint32_t l, uint32_t v, uint32_t *idx) {
    if (l == 0) return -3;
    if (v < a[0]) { *idx = 0xffffffffu; return -1; }
    if (v > a[l - 1]) { *idx = 0xffffffffu; return -2; }
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

constexpr uint32_t MMR_COV_SITES = 4096;  // windows with more sites report their read count as the coverage bound

__global__ void __launch_bounds__(RS_THREADS) methmer_size_kernel(MethmerParams P) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_base;
    __shared__ int s_diff[2][MMR_COV_SITES + 1];  // per direction: +1 where a read's site range begins, -1 behind its end
    __shared__ int s_maxcov;
    const uint32_t w = blockIdx.x;
    const WindowRec W = P.win[w];
    WindowState &S = P.state[w];
    const uint32_t n = S.n, n_sites = S.n_sites;
    const bool active = n > 0 && n_sites > 0 && S.status == 0;
    uint32_t running = 0;
    const uint32_t n_items = active ? n * 2 : 0;
    const bool count_cov = active && n_sites <= MMR_COV_SITES;
    if (count_cov) for (uint32_t i = threadIdx.x; i <= n_sites; i += RS_THREADS) { s_diff[0][i] = 0; s_diff[1][i] = 0; }
    if (threadIdx.x == 0) s_maxcov = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_items; base += RS_THREADS) {
        uint32_t it = base + threadIdx.x;
        uint32_t cap = 0, slot = 0, d = 0;
        if (it < n_items) {
            d = it >= n ? 1u : 0u;
            uint32_t id = it - d * n;
            slot = W.first_read + id;
            const uint32_t src = P.rs_src[slot];
            const uint32_t nc = P.r_ncalls[src];
            const uint32_t *cp = P.calls_pos + P.reads[src].calls_off;
            const uint32_t *starts = P.site_start[d] + W.site_off;
            uint32_t xl = 0, xr = 0;
            bool none = nc == 0;
            if (!none) {
                int st = search_left(starts, n_sites, cp[0], &xl);
                if (st == -2 || st == -3) none = true;
                else {
                    if (st == 0) xl = xl == 0 ? 0 : xl - 1;
                    st = search_left(starts, n_sites, cp[nc - 1], &xr);
                    if (st == -1 || st == -3) none = true;
                    else {
                        if (xl == 0xffffffffu) xl = 0;
                        if (xr == 0xffffffffu) xr = n_sites;
                    }
                }
            }
            if (none || xr <= xl) { xl = 0; xr = 0; cap = 0; }
            else {
                cap = (xr - xl) + 2u * (uint32_t)P.k + 2u;
                if (count_cov) {  // (methmers are counted at consecutive sites from the first complete one: at most `cap` of them)
                    const uint32_t xe = xl + cap < n_sites ? xl + cap : n_sites;
                    atomicAdd(&s_diff[d][xl], 1); atomicAdd(&s_diff[d][xe], -1);
                }
            }
            P.mm_xl[d][slot] = xl;
            P.mm_xr[d][slot] = xr;
        }
        uint32_t tot;
        uint32_t ex = block_exclusive_scan(cap, &tot, s_warp);
        if (it < n_items) P.mm_off[d][slot] = running + ex;  // window-local for now
        running += tot;
    }
    // the largest number of reads over one site (either direction): every count of the join tables stays below it
    if (count_cov) {
        __syncthreads();
        const uint32_t per = (n_sites + RS_THREADS - 1) / RS_THREADS;
        for (int d = 0; d < 2; d++) {
            const uint32_t a = threadIdx.x * per, z = a + per < n_sites ? a + per : n_sites;
            int sum = 0;
            for (uint32_t i = a; i < z; i++) sum += s_diff[d][i];
            uint32_t tot;
            int run = (int)block_exclusive_scan((uint32_t)sum, &tot, s_warp), mx = 0;  // (two's complement sums)
            for (uint32_t i = a; i < z; i++) { run += s_diff[d][i]; mx = run > mx ? run : mx; }
            atomicMax(&s_maxcov, mx);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        S.max_cov = count_cov ? (uint32_t)s_maxcov : n;
        s_base = atomicAdd(&P.pool_total[0], running);
        S.mmr_base[0] = s_base;
        S.mmr_base[1] = s_base;
        uint32_t tb = atomicAdd(&P.pool_total[1], active ? 2 * n_sites : 0);
        S.tab_base[0] = tb;
        S.tab_base[1] = tb + n_sites;
        if (active) atomicMax(&P.pool_total[2], n_sites);
    }
    __syncthreads();
    const uint32_t gbase = s_base;
    for (uint32_t it = threadIdx.x; it < n_items; it += RS_THREADS) {
        uint32_t d = it >= n ? 1u : 0u;
        uint32_t slot = W.first_read + (it - d * n);
        P.mm_off[d][slot] += gbase;
    }
}

constexpr int MMR_WARPS = 4;

// Persistent warps over a queue of (record, direction) items, longest records first (see decode_kernel).
__device__ void methmer_fill_item(const MethmerParams &P, uint32_t slot, uint32_t d);

__global__ void __launch_bounds__(MMR_WARPS * 32) methmer_fill_kernel(MethmerParams P) {
    const unsigned lane = lane_id();
    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(P.pool_total + 3, 1u);
        q = __shfl_sync(FULL_MASK, q, 0);
        if (q >= P.n_slots * 2) return;
        const uint32_t slot = P.order ? P.order[q >> 1] : q >> 1;
        methmer_fill_item(P, slot, q & 1u);
        __syncwarp();
    }
}

__device__ void methmer_fill_item(const MethmerParams &P, uint32_t slot, uint32_t d) {
    const unsigned lane = lane_id();
    const uint32_t w = P.read_win[slot];
    if (w == 0xffffffffu) return;  // record outside every window
    const WindowRec W = P.win[w];
    const WindowState &S = P.state[w];
    const uint32_t id = slot - W.first_read;
    if (id >= S.n || S.n_sites == 0 || S.status != 0) {
        if (lane == 0) { P.mm_n[d][slot] = 0; P.mm_start[d][slot] = 0; }
        return;
    }
    const uint32_t n_sites = S.n_sites;
    const uint32_t xl = P.mm_xl[d][slot], xr = P.mm_xr[d][slot];
    if (xr <= xl) {
        if (lane == 0) { P.mm_n[d][slot] = 0; P.mm_start[d][slot] = 0; }
        return;
    }
    const uint32_t src = P.rs_src[slot];
    const uint32_t nc = P.r_ncalls[src];
    const bool sorted = !(P.r_status[src] & RS_UNSORTED);
    const uint32_t *cp = P.calls_pos + P.reads[src].calls_off;
    const uint8_t *cc = P.calls_cat + P.reads[src].calls_off;
    const uint32_t *starts = P.site_start[d] + W.site_off;
    const uint8_t *lens = P.site_len[d] + W.site_off;
    const uint32_t off = P.mm_off[d][slot];
    const uint32_t cap = (xr - xl) + 2u * (uint32_t)P.k + 2u;
    if (off + cap > P.pool_cap) {  // engine sized the pool from pool_total: cannot happen
        if (lane == 0) { P.mm_n[d][slot] = 0; P.mm_start[d][slot] = 0; }
        return;
    }
    uint32_t *ent = P.ent_pool + off;
    uint32_t *out = P.mmr_pool + off;

    // ---- phase 1: entries (unique starts, except the i>1 rule) and their symbols ----
    uint32_t nE = 0;
    for (uint32_t base = xl; base < xr; base += 32) {
        const uint32_t i = base + lane;
        bool valid = i < xr;
        uint32_t st = 0;
        if (valid) {
            st = starts[i];
            if (i > 1 && starts[i - 1] == st) valid = false;  // blockjoin.c:3391
        }
        unsigned vm = __ballot_sync(FULL_MASK, valid);
        if (valid) {
            uint32_t sym = 2;  // '-'
            const bool shadowed = i == 0 && xr > 1 && starts[1] == st;  // next buffer entry is site 1, not a call
            if (!shadowed) {
                if (sorted) {
                    uint32_t lo = 0, hi = nc;
                    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (cp[mid] < st) lo = mid + 1; else hi = mid; }
                    if (lo < nc && cp[lo] == st) sym = cc[lo];
                } else {
                    for (uint32_t j = 0; j < nc; j++) if (cp[j] == st && cc[j] < sym) sym = cc[j];
                    // a call equal in position always sorts right after the site entry; the smallest
                    // category wins because the category is part of the sort key (blockjoin.c:3398)
                    bool any = false;
                    for (uint32_t j = 0; j < nc; j++) if (cp[j] == st) any = true;
                    if (!any) sym = 2;
                }
            }
            ent[nE + __popc(vm & ((1u << lane) - 1u))] = (i << 2) | (sym & 3u);
        }
        nE += __popc(vm);
    }
    __syncwarp();
    // ---- phase 2: keys ----
    uint32_t n_out = 0, first_j = 0xffffffffu;
    for (uint32_t ebase = 0; ebase < nE; ebase += 32) {
        const uint32_t e = ebase + lane;
        uint32_t cnt = 0, my_first = 0xffffffffu;
        uint32_t i0 = 0, st = 0;
        if (e < nE) {
            i0 = ent[e] >> 2;
            st = starts[i0];
            for (uint32_t j = i0; j < n_sites && starts[j] == st; j++) {
                uint32_t L = lens[j];
                if (e + L <= nE) { if (!cnt) my_first = j; cnt++; }
            }
        }
        uint32_t incl = warp_inclusive_sum(cnt);
        uint32_t o = n_out + incl - cnt;
        if (cnt) {
            for (uint32_t j = i0; j < n_sites && starts[j] == st; j++) {
                uint32_t L = lens[j];
                if (e + L <= nE) {
                    uint32_t key = 0;
                    for (uint32_t t = 0; t < L; t++) key = key << 2 | (ent[e + t] & 3u);
                    if (o < cap) out[o] = key;
                    o++;
                }
            }
        }
        unsigned hm = __ballot_sync(FULL_MASK, cnt != 0);
        if (first_j == 0xffffffffu && hm) first_j = __shfl_sync(FULL_MASK, my_first, __ffs(hm) - 1);
        n_out += __shfl_sync(FULL_MASK, incl, 31);
    }
    if (lane == 0) {
        P.mm_n[d][slot] = n_out;
        P.mm_start[d][slot] = n_out ? first_j : 0;  // store_mmr_of_one_read, blockjoin.c:3518-3529
    }
}

}  // namespace pomfret_gpu
#endif
