// (a) Staging without a host-side copy: when the caller's record buffers are registered (pinned and mapped,
// pomfret_gpu_host_register), the device gathers the record payloads itself.  One warp per record reads the
// fields (CIGAR, SEQ, MM, ML, MD) straight out of host memory over PCIe and lays them out in the batch blob:
// every field on a 16-byte boundary, zero padded — the layout add_reads() produces on the host otherwise.
// Host memory is read once, by the copy that has to happen anyway; the CPU touches only descriptors.
#ifndef POMFRET_GPU_GATHER_CUH
#define POMFRET_GPU_GATHER_CUH
#include "gpu_rt.h"
#include "types.h"

namespace pomfret_gpu {

constexpr int GATHER_WARPS = 8;

struct GatherSrc {
    uint64_t ptr[5];  // device-visible addresses of cigar, seq, mm, ml, md in the caller's memory; 0: not gathered
};

struct GatherParams {
    const ReadRec *reads;
    const GatherSrc *src;
    uint8_t *blob;
    uint32_t n_reads;
};

// 16 output bytes = bytes [sh, sh + 16) of the 32-byte window (a, b); sh = 4 * Q + r with Q a compile-time
// word offset and r the byte offset inside a word
template <int Q> __device__ __forceinline__ uint4 shift_window(const uint4 &a, const uint4 &b, uint32_t r8) {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 o;
    o.x = __funnelshift_r(w[Q + 0], w[Q + 1], r8);
    o.y = __funnelshift_r(w[Q + 1], w[Q + 2], r8);
    o.z = __funnelshift_r(w[Q + 2], w[Q + 3], r8);
    o.w = __funnelshift_r(w[Q + 3], w[Q + 4], r8);
    return o;
}

// dst is 16-byte aligned and owns align16(n) bytes; src has any alignment.  The source is read in 16-byte
// aligned chunks, lane-consecutive (512-byte requests on the bus, GATHER_U of them in flight per warp), and
// every output chunk is cut out of two neighbouring source chunks.
constexpr int GATHER_U = 4;

__device__ __forceinline__ void warp_gather_field(uint8_t *dst, uint64_t src_addr, uint32_t n) {
    const unsigned lane = lane_id();
    const uint32_t sh = (uint32_t)(src_addr & 15u);
    const uint4 *sc = reinterpret_cast<const uint4 *>(src_addr - sh);
    const uint32_t n_src = (n + sh + 15u) >> 4;  // aligned source chunks that hold the field
    const uint32_t n_out = (n + 15u) >> 4;       // chunks written, padding included
    uint4 *dc = reinterpret_cast<uint4 *>(dst);
    const uint32_t q = sh >> 2, r8 = (sh & 3u) * 8u;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (uint32_t c0 = 0; c0 < n_out; c0 += 32 * GATHER_U) {
        uint4 cur[GATHER_U], nxt[GATHER_U];
#pragma unroll
        for (int u = 0; u < GATHER_U; u++) {
            const uint32_t c = c0 + u * 32 + lane;
            cur[u] = c < n_src ? sc[c] : zero;
        }
#pragma unroll
        for (int u = 0; u < GATHER_U; u++) {
            // the chunk behind this lane's: the next lane's, or for lane 31 the first of the next group of 32
            nxt[u].x = __shfl_down_sync(FULL_MASK, cur[u].x, 1); nxt[u].y = __shfl_down_sync(FULL_MASK, cur[u].y, 1);
            nxt[u].z = __shfl_down_sync(FULL_MASK, cur[u].z, 1); nxt[u].w = __shfl_down_sync(FULL_MASK, cur[u].w, 1);
            uint4 f = zero;
            if (u + 1 < GATHER_U) {
                const uint4 &g = cur[u + 1 < GATHER_U ? u + 1 : u];
                f.x = __shfl_sync(FULL_MASK, g.x, 0); f.y = __shfl_sync(FULL_MASK, g.y, 0);
                f.z = __shfl_sync(FULL_MASK, g.z, 0); f.w = __shfl_sync(FULL_MASK, g.w, 0);
            } else if (lane == 31 && sh) {
                const uint32_t c = c0 + GATHER_U * 32;
                if (c < n_src) f = sc[c];
            }
            if (lane == 31) nxt[u] = f;
        }
#pragma unroll
        for (int u = 0; u < GATHER_U; u++) {
            const uint32_t c = c0 + u * 32 + lane;
            uint4 o;
            if (sh == 0) o = cur[u];
            else if (q == 0) o = shift_window<0>(cur[u], nxt[u], r8);
            else if (q == 1) o = shift_window<1>(cur[u], nxt[u], r8);
            else if (q == 2) o = shift_window<2>(cur[u], nxt[u], r8);
            else o = shift_window<3>(cur[u], nxt[u], r8);
            // bytes at and behind the end of the field are padding
            const uint32_t b0 = c * 16u;
            if (b0 + 16u > n) {
                uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t bw = b0 + 4u * i;
                    if (bw >= n) w[i] = 0u;
                    else if (n - bw < 4u) w[i] &= (1u << ((n - bw) * 8u)) - 1u;
                }
                o = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (c < n_out) dc[c] = o;
        }
    }
}

__global__ void __launch_bounds__(GATHER_WARPS * 32) gather_kernel(GatherParams P) {
    const uint32_t ri = blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5);
    if (ri >= P.n_reads) return;
    const ReadRec &R = P.reads[ri];
    const GatherSrc &S = P.src[ri];
    if (S.ptr[0]) warp_gather_field(P.blob + (size_t)R.cigar_off * 16, S.ptr[0], R.n_cigar * 4u);
    if (S.ptr[1]) {
        warp_gather_field(P.blob + (size_t)R.seq_off * 16, S.ptr[1], (R.l_qseq + 1u) >> 1);
        __syncwarp();
        // the unused low nibble of an odd-length SEQ must read as "no base"
        if ((R.l_qseq & 1u) && lane_id() == 0) P.blob[(size_t)R.seq_off * 16 + (R.l_qseq >> 1)] &= 0xf0u;
    }
    if (S.ptr[2]) warp_gather_field(P.blob + (size_t)R.mm_off * 16, S.ptr[2], R.mm_len);
    if (S.ptr[3]) warp_gather_field(P.blob + (size_t)R.ml_off * 16, S.ptr[3], R.ml_len);
    if (S.ptr[4]) warp_gather_field(P.blob + (size_t)R.md_off * 16, S.ptr[4], R.md_len);
}

}  // namespace pomfret_gpu
#endif
