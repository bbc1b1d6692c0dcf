// Runtime glue for the sm_100a engine: CUDA runtime include, launch macro, error checks,
// warp helpers.  The only non-CUDA branch is the test-only functional emulator hook
// (tests/cuda_emu/, never part of libpomfret_gpu.so).
#ifndef POMFRET_GPU_RT_H
#define POMFRET_GPU_RT_H

#ifdef POMFRET_CUDA_EMU
#include "cuda_emu.h"
#define POMFRET_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cuda_emu::launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })
#define POMFRET_DYN_SMEM(type, name) type *name = (type *)cuda_emu::dyn_smem()
#else
#include <cuda_runtime.h>
#define POMFRET_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define POMFRET_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; \
    type *name = reinterpret_cast<type *>(name##_raw_)
#endif

#include <stdint.h>

#define FULL_MASK 0xffffffffu

namespace pomfret_gpu {

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// inclusive warp scan (sum)
template <typename T> __device__ __forceinline__ T warp_inclusive_sum(T v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(FULL_MASK, v, o);
        if (lane_id() >= (unsigned)o) v += t;
    }
    return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T t = __shfl_xor_sync(FULL_MASK, v, o);
        v = t > v ? t : v;
    }
    return v;
}
template <typename T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T t = __shfl_xor_sync(FULL_MASK, v, o);
        v = t < v ? t : v;
    }
    return v;
}

}  // namespace pomfret_gpu
#endif
