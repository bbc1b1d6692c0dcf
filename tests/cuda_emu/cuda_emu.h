// TEST INFRASTRUCTURE — a tiny functional SIMT emulator, NOT a product code path.
//
// There is no GPU in the build container, and GPU-box minutes are scarce.  This header lets the
// *unmodified* kernel sources of pomfret_b200/csrc/gpu be compiled with g++ and stepped on the CPU
// so that their logic (warp collectives, barriers, shared memory, atomics) can be debugged and
// checked against the oracle before they ever run on a B200.  Each CUDA thread is a ucontext
// fiber; fibers of one block are scheduled round-robin and switch only at synchronisation points
// (__syncthreads, __syncwarp, shuffles, votes), blocks run one after another.  A collective that
// not every expected lane reaches is reported as a deadlock, which is how divergence bugs show up.
//
// The emulated library is built into tests/cuda_emu/_build/ by tests/cuda_emu/build_emu.py and is
// only ever loaded by tests marked `emu`.  libpomfret_gpu.so (the product) is always the nvcc
// build for sm_100a and refuses to initialise without a CUDA device.
#ifndef POMFRET_CUDA_EMU_H
#define POMFRET_CUDA_EMU_H
#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
struct uint2 { uint32_t x, y; } __attribute__((aligned(8)));
struct int2 { int x, y; } __attribute__((aligned(8)));
struct ushort2 { uint16_t x, y; };
struct float2 { float x, y; } __attribute__((aligned(8)));
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace cuda_emu {
struct ThreadCtx {
    dim3 tid, bid, bdim, gdim;
};
ThreadCtx &ctx();
void sync_block();
void sync_warp(unsigned mask);
uint64_t warp_exchange(unsigned mask, uint64_t value, int src_lane);       // value of src_lane
unsigned warp_ballot(unsigned mask, int pred);
unsigned warp_match_any(unsigned mask, uint64_t value);
void *dyn_smem();
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
}  // namespace cuda_emu

#define threadIdx (cuda_emu::ctx().tid)
#define blockIdx (cuda_emu::ctx().bid)
#define blockDim (cuda_emu::ctx().bdim)
#define gridDim (cuda_emu::ctx().gdim)
#define warpSize 32

static inline void __syncthreads() { cuda_emu::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cuda_emu::sync_warp(mask); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

namespace cuda_emu {
static inline int lane_id() { return (int)(ctx().tid.x & 31); }
template <typename T> static inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload too wide");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T> static inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
}  // namespace cuda_emu

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int lane = cuda_emu::lane_id();
    int s = (lane & ~(width - 1)) | (src & (width - 1));
    return cuda_emu::from_bits<T>(cuda_emu::warp_exchange(mask, cuda_emu::to_bits(v), s));
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = cuda_emu::lane_id();
    int s = lane - (int)delta;
    if (s < (lane & ~(width - 1))) s = lane;
    return cuda_emu::from_bits<T>(cuda_emu::warp_exchange(mask, cuda_emu::to_bits(v), s));
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = cuda_emu::lane_id();
    int s = lane + (int)delta;
    if (s > (lane | (width - 1))) s = lane;
    return cuda_emu::from_bits<T>(cuda_emu::warp_exchange(mask, cuda_emu::to_bits(v), s));
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    int lane = cuda_emu::lane_id();
    int s = lane ^ lane_mask;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return cuda_emu::from_bits<T>(cuda_emu::warp_exchange(mask, cuda_emu::to_bits(v), s));
}
static inline unsigned __ballot_sync(unsigned mask, int pred) { return cuda_emu::warp_ballot(mask, pred); }
static inline int __any_sync(unsigned mask, int pred) { return cuda_emu::warp_ballot(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return cuda_emu::warp_ballot(mask, !pred) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    return cuda_emu::warp_match_any(mask, cuda_emu::to_bits(v));
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) { unsigned t = __shfl_xor_sync(mask, v, o); v = t > v ? t : v; }
    return v;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) { unsigned t = __shfl_xor_sync(mask, v, o); v = t < v ? t : v; }
    return v;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(mask, v, o);
    return v;
}

// ---- integer / float intrinsics ----
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    // position of the offset-th set bit at or above `base` (offset > 0) / below (offset < 0)
    if (offset > 0) {
        for (unsigned p = base; p < 32; p++) if ((mask >> p) & 1u) { if (--offset == 0) return p; }
    } else if (offset < 0) {
        for (int p = (int)base; p >= 0; p--) if ((mask >> p) & 1u) { if (++offset == 0) return (unsigned)p; }
    } else if ((mask >> base) & 1u) return base;
    return 0xffffffffu;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (shift & 31u));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (shift & 31u)) >> 32);
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0;
        r |= b << (8 * i);
    }
    return r;
}
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __uint2float_rn(unsigned x) { return (float)x; }
static inline float __int2float_rn(int x) { return (float)x; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

// ---- atomics (shared and global memory are both ordinary host memory here) ----
template <typename T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicSub(T *p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicMax(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <typename T> static inline T atomicMin(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <typename T> static inline T atomicCAS(T *p, T cmp, T v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}

// ---- runtime API subset ----
typedef int cudaError_t;
typedef struct cuda_emu_stream *cudaStream_t;
typedef struct cuda_emu_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100, cudaErrorHostMemoryAlreadyRegistered = 712 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t totalGlobalMem; int major, minor; size_t sharedMemPerBlockOptin; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t n);
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void *p);
enum { cudaHostRegisterPortable = 1, cudaHostRegisterMapped = 2 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = nullptr; a->hostPointer = nullptr; return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void *p, int v, size_t n);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }  // launches run synchronously here
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char *cudaGetErrorString(cudaError_t e);
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

#endif
