// (a/d) Read-set construction, one CTA per window: everything load_reads_given_interval does after a
// record has been decoded (reference blockjoin.c:1111-1163): compaction of the kept records into read
// ids, the end-sorted `revbuf`, the left/right (strict) reference-read lists and the left-coverage gate.
#ifndef POMFRET_GPU_READSET_CUH
#define POMFRET_GPU_READSET_CUH
#include "gpu_rt.h"
#include "types.h"

namespace pomfret_gpu {

constexpr int RS_THREADS = 256;

struct ReadsetParams {
    const ReadRec *reads;
    const WindowRec *win;
    WindowState *state;
    const uint32_t *r_status, *r_end, *r_ncalls;
    int32_t *r_id;        // per batch read: id inside its window's read set, -1 if dropped
    uint32_t *rs_src;     // [first_read + id] -> batch read index
    uint32_t *rs_rev;     // [first_read + rank] -> id, ascending (end, id)
    uint32_t *ids_left, *ids_left_strict, *ids_right, *ids_right_strict;  // slices at first_read
    int32_t *rs_hp;       // [first_read + id] initial haplotag
    uint32_t n_windows;
};

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total gets the block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *s_warp /* [32] */) {
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    uint32_t incl = warp_inclusive_sum(v);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < n_warps ? s_warp[lane] : 0;
        uint32_t wi = warp_inclusive_sum(w);
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    uint32_t res = s_warp[warp] + incl - v;
    *total = s_warp[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(RS_THREADS) readset_kernel(ReadsetParams P) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_ends[1024];
    const uint32_t w = blockIdx.x;
    const WindowRec W = P.win[w];
    const uint32_t first = W.first_read, nr = W.n_reads;
    const uint32_t tid = threadIdx.x;
    if (tid < 8) s_cnt[tid] = 0;
    __syncthreads();
    // ---- ids: ordered compaction of kept records ----
    uint32_t n_loaded = 0, fatal = 0;
    for (uint32_t base = 0; base < nr; base += RS_THREADS) {
        uint32_t i = base + tid;
        uint32_t keep = 0;
        if (i < nr) {
            uint32_t st = P.r_status[first + i];
            keep = (st & RS_KEPT) ? 1u : 0u;
            if (st & RS_FATAL_CIGAR) fatal = 1;
        }
        uint32_t tot;
        uint32_t ex = block_exclusive_scan(keep, &tot, s_warp);
        if (i < nr) {
            if (keep) {
                uint32_t id = n_loaded + ex;
                P.r_id[first + i] = (int32_t)id;
                P.rs_src[first + id] = first + i;
                P.rs_hp[first + id] = P.reads[first + i].hp;
            } else P.r_id[first + i] = -1;
        }
        n_loaded += tot;
    }
    __syncthreads();
    // ---- left / right lists in id order, coverage check, call total ----
    const uint32_t itvl_s = W.ref_start, itvl_e = W.ref_end;
    uint32_t nL = 0, nLS = 0, nR = 0, nRS = 0;
    uint32_t cov0 = 0, cov1 = 0, calls = 0;
    for (uint32_t base = 0; base < n_loaded; base += RS_THREADS) {
        uint32_t id = base + tid;
        uint32_t fl = 0, fls = 0, fr = 0, frs = 0;
        if (id < n_loaded) {
            uint32_t src = P.rs_src[first + id];
            uint32_t start = P.reads[src].pos, end = P.r_end[src];
            int32_t hp = P.reads[src].hp;
            calls += P.r_ncalls[src];
            if (start <= itvl_s) {
                fl = 1;
                fls = end > itvl_s;
                if (hp == 0) cov0++;
                if (hp == 1) cov1++;
            } else if (end >= itvl_e) {
                fr = 1;
                frs = start < itvl_e;
            }
        }
        uint32_t tot, ex;
        ex = block_exclusive_scan(fl, &tot, s_warp);
        if (fl) P.ids_left[first + nL + ex] = id;
        nL += tot;
        ex = block_exclusive_scan(fls, &tot, s_warp);
        if (fls) P.ids_left_strict[first + nLS + ex] = id;
        nLS += tot;
        ex = block_exclusive_scan(fr, &tot, s_warp);
        if (fr) P.ids_right[first + nR + ex] = id;
        nR += tot;
        ex = block_exclusive_scan(frs, &tot, s_warp);
        if (frs) P.ids_right_strict[first + nRS + ex] = id;
        nRS += tot;
    }
    cov0 = warp_sum(cov0); cov1 = warp_sum(cov1); calls = warp_sum(calls);
    fatal = __any_sync(FULL_MASK, fatal);
    if (lane_id() == 0) {
        atomicAdd(&s_cnt[0], cov0);
        atomicAdd(&s_cnt[1], cov1);
        atomicAdd(&s_cnt[2], calls);
        if (fatal) atomicOr(&s_cnt[3], 1u);
    }
    __syncthreads();
    // ---- revbuf: rank of (end, id) among the kept reads (radix_sort_ksu64 of end<<32|id, :1126,1140) ----
    for (uint32_t base = 0; base < n_loaded; base += RS_THREADS) {
        uint32_t id = base + tid;
        uint32_t my_end = 0;
        if (id < n_loaded) my_end = P.r_end[P.rs_src[first + id]];
        uint32_t rank = 0;
        for (uint32_t tb = 0; tb < n_loaded; tb += 1024) {
            uint32_t tn = n_loaded - tb < 1024 ? n_loaded - tb : 1024;
            __syncthreads();
            for (uint32_t j = tid; j < tn; j += RS_THREADS) s_ends[j] = P.r_end[P.rs_src[first + tb + j]];
            __syncthreads();
            if (id < n_loaded)
                for (uint32_t j = 0; j < tn; j++) {
                    uint32_t e = s_ends[j], oid = tb + j;
                    rank += (e < my_end) || (e == my_end && oid < id);
                }
        }
        if (id < n_loaded) P.rs_rev[first + rank] = id;
    }
    if (tid == 0) {
        WindowState &S = P.state[w];
        S.n_loaded = n_loaded;
        S.n = (s_cnt[0] < 15 || s_cnt[1] < 15) ? 0 : n_loaded;  // HARD_COV_THRESHOLD, :1161
        S.n_left = nL; S.n_left_strict = nLS; S.n_right = nR; S.n_right_strict = nRS;
        S.total_calls = s_cnt[2];
        S.status = s_cnt[3] ? -6 : 0;
        S.n_sites = 0;
        S.mmr_total[0] = S.mmr_total[1] = 0;
        S.n_order[0] = S.n_order[1] = 0;
        for (int d = 0; d < 2; d++) for (int t = 0; t < 4; t++) S.table[d][t] = 0;
    }
}

}  // namespace pomfret_gpu
#endif
