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

// dst is 16-byte aligned and owns align16(n) bytes; src has any alignment.  Source words are read 4-byte
// aligned (lane-consecutive: 128-byte requests on the bus) and shifted into place.
__device__ __forceinline__ void warp_gather_field(uint8_t *dst, uint64_t src_addr, uint32_t n) {
    const unsigned lane = lane_id();
    const uint32_t sh = (uint32_t)(src_addr & 3u);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src_addr - sh);
    const uint32_t n_src_words = (n + sh + 3u) >> 2;     // aligned words that hold the field
    const uint32_t n_out_words = ((n + 15u) & ~15u) >> 2;  // words written, padding included
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst);
    constexpr int U = 8;  // 8 x 128 bytes per warp in flight: reads over PCIe are latency bound
    for (uint32_t w0 = 0; w0 < n_out_words; w0 += 32 * U) {
        uint32_t lo[U], nx[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t w = w0 + u * 32 + lane;
            lo[u] = w < n_src_words ? sw[w] : 0u;
        }
        // the word behind lane 31's: first word of the next group of 32
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t first_next = u + 1 < U ? __shfl_sync(FULL_MASK, lo[u + 1 < U ? u + 1 : u], 0) : 0u;
            nx[u] = __shfl_down_sync(FULL_MASK, lo[u], 1);
            if (lane == 31) {
                const uint32_t w = w0 + u * 32 + 32;
                nx[u] = u + 1 < U ? first_next : (w < n_src_words ? sw[w] : 0u);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t w = w0 + u * 32 + lane;
            uint32_t v = sh ? __funnelshift_r(lo[u], nx[u], sh * 8u) : lo[u];
            // bytes at and behind the end of the field are padding
            const uint32_t b0 = w * 4u;
            if (b0 >= n) v = 0u;
            else if (n - b0 < 4u) v &= (1u << ((n - b0) * 8u)) - 1u;
            if (w < n_out_words) dw[w] = v;
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
