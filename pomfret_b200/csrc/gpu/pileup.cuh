// (d) Per-CpG pileup and methmer-site layout.
//
// Replaces get_methmer_sites_and_ranges (reference blockjoin.c:3202-3354): per reference position the
// number of methylated / unmethylated calls over all reads of the window, selection of the positions
// with at least cov_for_selection of both, ascending order, and per site the methmer length and start
// for the forward and the backward direction.
//
// pileup_tile_kernel: one CTA per (window, position tile).  The tile's counters live in shared memory
// as one packed 32-bit word per position (low half methylated, high half unmethylated).  One record per
// thread first: does it touch the tile, and which of its (sorted) calls fall into it (two binary
// searches); the records that do are listed in shared memory and then streamed one per warp step with
// lane-consecutive loads into the counters (shared-memory atomics); no-call entries are not needed for
// selection and are skipped.  The
// strand-saturation bits of the reference counter never reach an output and are not materialised; its
// count field is 12 bits wide (u16 >> 4), which is reproduced by masking with 0xfff.
// sites_finalize_kernel: one CTA per window, concatenates the tile outputs in position order and
// derives (start, length) of the methmer anchored at every site for both directions.
#ifndef POMFRET_GPU_PILEUP_CUH
#define POMFRET_GPU_PILEUP_CUH
#include "gpu_rt.h"
#include "types.h"
#include "readset.cuh"

namespace pomfret_gpu {

constexpr int PILE_THREADS = 512;
constexpr uint32_t PILE_TILE = 24 * 1024;  // positions per tile: 96 KB of counters, two CTAs per SM

struct TileRec {
    uint32_t window;
    uint32_t tile;      // index inside the window
    uint32_t out_off;   // slice of the tile-output array
    uint32_t out_cap;
};

struct PileupParams {
    const WindowRec *win;
    const WindowState *state;
    const TileRec *tiles;
    const uint32_t *win_base;     // per window: position that maps to counter 0 (min read start - 1)
    const ReadRec *reads;
    const uint32_t *rs_src;
    const uint32_t *r_ncalls, *r_status, *r_end;
    const uint32_t *calls_pos;
    const uint8_t *calls_cat;
    uint32_t *tile_out;           // selected positions per tile, ascending
    uint32_t *tile_count;
    uint32_t cov;
};

__global__ void __launch_bounds__(PILE_THREADS) pileup_tile_kernel(PileupParams P) {
    POMFRET_DYN_SMEM(uint32_t, cnt);  // PILE_TILE packed counters
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_hits, s_hit_off[PILE_THREADS], s_hit_a[PILE_THREADS], s_hit_e[PILE_THREADS];
    const TileRec T = P.tiles[blockIdx.x];
    const WindowRec W = P.win[T.window];
    const uint32_t n = P.state[T.window].n;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5, n_warps = PILE_THREADS / 32;
    const uint32_t wbase = P.win_base[T.window];
    const uint32_t t0 = T.tile * PILE_TILE;  // window-relative
    for (uint32_t i = tid; i < PILE_TILE; i += PILE_THREADS) cnt[i] = 0;
    __syncthreads();
    // Positions are handled relative to the window base in wrapping 32-bit arithmetic, compared as signed:
    // the reference's position arithmetic wraps too, and a call that a clipped record
    // maps left of the base stays ordered (negative) instead of turning into a huge offset.
    // Step 1, one record per thread: metadata, "does the record touch this tile" (most do not) and the range
    // of its calls inside the tile — up to 512 independent chains of dependent loads at a time; the records
    // that touch the tile are listed in shared memory.  Step 2, one listed record per warp step: its calls are
    // streamed with lane-consecutive loads (known trip count) into the shared counters.
    for (uint32_t id0 = 0; id0 < n; id0 += PILE_THREADS) {
        if (tid == 0) s_hits = 0;
        __syncthreads();
        const uint32_t id = id0 + tid;
        if (id < n) {
            const uint32_t src = P.rs_src[W.first_read + id];
            const uint32_t nc = P.r_ncalls[src];
            if (nc) {
                const uint32_t coff = P.reads[src].calls_off;
                const uint32_t *cp = P.calls_pos + coff;
                uint32_t a = 0, e = nc;
                bool hit = true;
                if (!(P.r_status[src] & RS_UNSORTED)) {
                    hit = !((int32_t)(cp[nc - 1] - wbase) < (int32_t)t0 || (int32_t)(cp[0] - wbase) >= (int32_t)(t0 + PILE_TILE));
                    if (hit) {
                        uint32_t lo = 0, hi = nc;
                        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if ((int32_t)(cp[mid] - wbase) < (int32_t)t0) lo = mid + 1; else hi = mid; }
                        a = lo;
                        hi = nc;
                        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if ((int32_t)(cp[mid] - wbase) < (int32_t)(t0 + PILE_TILE)) lo = mid + 1; else hi = mid; }
                        e = lo;
                        hit = e > a;
                    }
                }
                if (hit) {
                    const uint32_t k = atomicAdd(&s_hits, 1u);  // order is irrelevant: the counters only add
                    s_hit_off[k] = coff; s_hit_a[k] = a; s_hit_e[k] = e;
                }
            }
        }
        __syncthreads();
        const uint32_t nh = s_hits;
        for (uint32_t h = warp; h < nh; h += n_warps) {
            const uint32_t *cp = P.calls_pos + s_hit_off[h];
            const uint8_t *cc = P.calls_cat + s_hit_off[h];
            const uint32_t e = s_hit_e[h];
            for (uint32_t j = s_hit_a[h] + lane; j < e; j += 32) {
                const uint32_t rel = cp[j] - wbase - t0;
                const uint32_t cat = cc[j];
                if (rel < PILE_TILE && cat < 2u) atomicAdd(&cnt[rel], cat == 0u ? 1u : 0x10000u);
            }
        }
        __syncthreads();
    }
    __syncthreads();
    // selection + ordered compaction
    constexpr uint32_t PER = PILE_TILE / PILE_THREADS;
    const uint32_t p0 = tid * PER;
    uint32_t mine = 0;
    for (uint32_t i = 0; i < PER; i++) {
        uint32_t c = cnt[p0 + i];
        mine += ((c & 0xfffu) >= P.cov && ((c >> 16) & 0xfffu) >= P.cov) ? 1u : 0u;
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(mine, &total, s_warp);
    if (mine) {
        uint32_t o = ex;
        for (uint32_t i = 0; i < PER; i++) {
            uint32_t c = cnt[p0 + i];
            if ((c & 0xfffu) >= P.cov && ((c >> 16) & 0xfffu) >= P.cov) {
                if (o < T.out_cap) P.tile_out[T.out_off + o] = wbase + t0 + p0 + i;
                o++;
            }
        }
    }
    if (tid == 0) P.tile_count[blockIdx.x] = total;
}

struct SitesParams {
    const WindowRec *win;
    WindowState *state;
    const TileRec *tiles;
    const uint32_t *win_tile_first;  // per window: index of its first TileRec; [n_windows] = total
    const uint32_t *tile_out, *tile_count;
    uint32_t *site_pos;
    uint32_t *site_start[2];
    uint8_t *site_len[2];
    int32_t k, k_span;
};

__global__ void __launch_bounds__(256) sites_finalize_kernel(SitesParams P) {
    __shared__ uint32_t s_n;
    const uint32_t w = blockIdx.x;
    const WindowRec W = P.win[w];
    const uint32_t tf = P.win_tile_first[w], tl = P.win_tile_first[w + 1];
    // concatenate (tiles are few: sequential offsets, parallel copies)
    uint32_t off = 0;
    bool overflow = false;
    for (uint32_t t = tf; t < tl; t++) {
        uint32_t c = P.tile_count[t];
        if (c > P.tiles[t].out_cap || off + c > W.site_cap) { overflow = true; break; }
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) P.site_pos[W.site_off + off + i] = P.tile_out[P.tiles[t].out_off + i];
        off += c;
    }
    if (threadIdx.x == 0) {
        s_n = overflow ? 0 : off;
        P.state[w].n_sites = s_n;
        if (overflow) P.state[w].status = -8;
    }
    __syncthreads();
    const int n = (int)s_n;
    const uint32_t *pos = P.site_pos + W.site_off;
    const uint32_t span = (uint32_t)P.k_span;
    // blockjoin.c:3307-3329
    for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) {
        // forward: j = min(i+k, n-1); shrink while the span is too long
        int j = i + P.k;
        if (j > n - 1) j = n - 1;
        while (pos[j] - pos[i] > span) j--;
        P.site_len[0][W.site_off + i] = (uint8_t)(j - i == 0 ? 1 : j - i);
        P.site_start[0][W.site_off + i] = pos[i];
        // backward: the same loop on the reversed array, expressed in ascending indices
        int jb = i - P.k;
        if (jb < 0) jb = 0;
        while (pos[i] - pos[jb] > span) jb++;
        P.site_len[1][W.site_off + i] = (uint8_t)(i - jb == 0 ? 1 : i - jb);
        P.site_start[1][W.site_off + i] = pos[jb];
    }
}

}  // namespace pomfret_gpu
#endif
