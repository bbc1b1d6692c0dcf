// TEST INFRASTRUCTURE — fiber scheduler and runtime stubs of the SIMT emulator (see cuda_emu.h).
#include "cuda_emu.h"
#include <ucontext.h>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <vector>

namespace cuda_emu {

namespace {

constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
    ucontext_t uc;
    char *stack = nullptr;
    ThreadCtx tc;
    bool done = false;
    int wait_kind = 0;  // 0 runnable, 1 block barrier, 2 warp barrier
    uint64_t wait_gen = 0;
    int warp = 0, lane = 0;
};

struct WarpState {
    uint64_t gen = 0;
    unsigned arrived = 0;
    unsigned pending_mask = 0;
    unsigned alive = 0;
    uint64_t slots[2][32];
    unsigned ballot[2] = {0, 0};
};

struct BlockState {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    uint64_t block_gen = 0;
    int block_arrived = 0;
    int n_alive = 0;
    ucontext_t sched;
    int cur = -1;
    void *dyn = nullptr;
    const std::function<void()> *body = nullptr;
};

BlockState *g_blk = nullptr;
std::mutex g_launch_mutex;
std::vector<char *> g_stack_pool;

void yield_to_scheduler() {
    Fiber &f = g_blk->fibers[g_blk->cur];
    swapcontext(&f.uc, &g_blk->sched);
}

void fiber_entry() {
    (*g_blk->body)();
    Fiber &f = g_blk->fibers[g_blk->cur];
    f.done = true;
    swapcontext(&f.uc, &g_blk->sched);
}

void try_release_warp(WarpState &w) {
    unsigned expected = w.pending_mask & w.alive;
    if (w.arrived && (w.arrived & expected) == expected) {
        w.gen++;
        w.arrived = 0;
        w.pending_mask = 0;
        w.ballot[w.gen & 1] = 0;
    }
}

void try_release_block(BlockState &b) {
    if (b.block_arrived > 0 && b.block_arrived >= b.n_alive) {
        b.block_gen++;
        b.block_arrived = 0;
    }
}

void warp_arrive_and_wait(unsigned mask) {
    Fiber &f = g_blk->fibers[g_blk->cur];
    WarpState &w = g_blk->warps[f.warp];
    if (!(mask >> f.lane & 1u)) {
        fprintf(stderr, "[cuda_emu] lane %d of warp %d called a collective with mask %08x that excludes it\n", f.lane, f.warp, mask);
        abort();
    }
    if (w.arrived == 0) w.pending_mask = mask;
    else if (w.pending_mask != mask) {
        fprintf(stderr, "[cuda_emu] warp %d: lanes disagree on collective mask (%08x vs %08x)\n", f.warp, w.pending_mask, mask);
        abort();
    }
    uint64_t my_gen = w.gen;
    w.arrived |= 1u << f.lane;
    try_release_warp(w);
    if (w.gen != my_gen) return;
    f.wait_kind = 2;
    f.wait_gen = my_gen;
    yield_to_scheduler();
}

}  // namespace

ThreadCtx &ctx() { return g_blk->fibers[g_blk->cur].tc; }
void *dyn_smem() { return g_blk->dyn; }

void sync_block() {
    BlockState &b = *g_blk;
    Fiber &f = b.fibers[b.cur];
    uint64_t my_gen = b.block_gen;
    b.block_arrived++;
    try_release_block(b);
    if (b.block_gen != my_gen) return;
    f.wait_kind = 1;
    f.wait_gen = my_gen;
    yield_to_scheduler();
}

void sync_warp(unsigned mask) { warp_arrive_and_wait(mask); }

uint64_t warp_exchange(unsigned mask, uint64_t value, int src_lane) {
    Fiber &f = g_blk->fibers[g_blk->cur];
    WarpState &w = g_blk->warps[f.warp];
    int par = (int)(w.gen & 1);
    w.slots[par][f.lane] = value;
    warp_arrive_and_wait(mask);
    // (a source lane that already left the kernel still has its slot: it wrote before arriving)
    if (src_lane < 0 || src_lane > 31 || !(mask >> src_lane & 1u)) return value;
    return w.slots[par][src_lane];
}

unsigned warp_ballot(unsigned mask, int pred) {
    Fiber &f = g_blk->fibers[g_blk->cur];
    WarpState &w = g_blk->warps[f.warp];
    int par = (int)(w.gen & 1);
    if (pred) w.ballot[par] |= 1u << f.lane;
    warp_arrive_and_wait(mask);
    return w.ballot[par] & mask;
}

unsigned warp_match_any(unsigned mask, uint64_t value) {
    Fiber &f = g_blk->fibers[g_blk->cur];
    WarpState &w = g_blk->warps[f.warp];
    int par = (int)(w.gen & 1);
    w.slots[par][f.lane] = value;
    warp_arrive_and_wait(mask);
    unsigned m = 0, part = mask;
    for (int l = 0; l < 32; l++)
        if ((part >> l & 1u) && w.slots[par][l] == value) m |= 1u << l;
    return m;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    const int n_threads = (int)(block.x * block.y * block.z);
    if (n_threads <= 0 || n_threads > 1024) { fprintf(stderr, "[cuda_emu] bad block size %d\n", n_threads); abort(); }
    BlockState blk;
    blk.fibers.resize(n_threads);
    blk.warps.resize((n_threads + 31) / 32);
    blk.body = &body;
    while ((int)g_stack_pool.size() < n_threads) g_stack_pool.push_back((char *)malloc(kStackBytes));
    std::vector<uint8_t> dyn(smem_bytes + 64);
    blk.dyn = (void *)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    g_blk = &blk;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                memset(dyn.data(), 0xCD, dyn.size());
                blk.block_gen = 0;
                blk.block_arrived = 0;
                blk.n_alive = n_threads;
                for (auto &w : blk.warps) { w = WarpState(); }
                for (int t = 0; t < n_threads; t++) {
                    Fiber &f = blk.fibers[t];
                    f.stack = g_stack_pool[t];
                    f.done = false;
                    f.wait_kind = 0;
                    f.warp = t / 32;
                    f.lane = t % 32;
                    blk.warps[f.warp].alive |= 1u << f.lane;
                    f.tc.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    f.tc.bid = dim3(bx, by, bz);
                    f.tc.bdim = block;
                    f.tc.gdim = grid;
                    getcontext(&f.uc);
                    f.uc.uc_stack.ss_sp = f.stack;
                    f.uc.uc_stack.ss_size = kStackBytes;
                    f.uc.uc_link = &blk.sched;
                    makecontext(&f.uc, (void (*)())fiber_entry, 0);
                }
                while (blk.n_alive > 0) {
                    bool progressed = false;
                    for (int t = 0; t < n_threads; t++) {
                        Fiber &f = blk.fibers[t];
                        if (f.done) continue;
                        if (f.wait_kind == 1 && blk.block_gen == f.wait_gen) continue;
                        if (f.wait_kind == 2 && blk.warps[f.warp].gen == f.wait_gen) continue;
                        f.wait_kind = 0;
                        blk.cur = t;
                        swapcontext(&blk.sched, &f.uc);
                        progressed = true;
                        if (f.done) {
                            blk.n_alive--;
                            WarpState &w = blk.warps[f.warp];
                            w.alive &= ~(1u << f.lane);
                            try_release_warp(w);
                            try_release_block(blk);
                        }
                    }
                    if (!progressed) {
                        fprintf(stderr, "[cuda_emu] DEADLOCK in block (%u,%u,%u): %d threads alive, none runnable\n", bx, by, bz, blk.n_alive);
                        for (int t = 0; t < n_threads; t++) {
                            Fiber &f = blk.fibers[t];
                            if (!f.done) fprintf(stderr, "  thread %d waits on %s\n", t, f.wait_kind == 1 ? "__syncthreads" : "warp collective");
                        }
                        abort();
                    }
                }
            }
    g_blk = nullptr;
}

}  // namespace cuda_emu

// ---------------- runtime stubs ----------------
struct cuda_emu_stream { int dummy; };
struct cuda_emu_event { std::chrono::steady_clock::time_point t; };

cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "cuda_emu (CPU functional emulator, tests only)");
    p->multiProcessorCount = 148;
    p->totalGlobalMem = (size_t)8 << 30;
    p->major = 10;
    p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t n) {
    *p = nullptr;
    if (posix_memalign(p, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
    memset(*p, 0xAB, n);
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { if (n) memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(dst, src, n, k); }
cudaError_t cudaMemset(void *p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new cuda_emu_stream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cuda_emu_event(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuda_emu error"; }
