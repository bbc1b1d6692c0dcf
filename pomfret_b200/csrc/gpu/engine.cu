// libpomfret_gpu: batch staging, kernel launches and the C ABI declared in include/pomfret_gpu.h.
//
// One batch = one region chunk of one host worker: packed records in a pinned arena, one async H2D
// copy, then decode -> read sets -> pileup/sites -> methmers -> greedy join on the batch's stream.
// The only host round trips are (1) the pool-size read after methmer sizing (two integers) and
// (2) collect().  There is no CPU fallback: without a CUDA device init() fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <algorithm>
#if defined(__x86_64__) && !defined(POMFRET_CUDA_EMU)
#include <emmintrin.h>
#define POMFRET_NT_COPY 1
#endif
#include "gpu_rt.h"
#include "types.h"
#include "decode.cuh"
#include "readset.cuh"
#include "pileup.cuh"
#include "methmer.cuh"
#include "join.cuh"
#include "haptag.cuh"
#include "gather.cuh"
#include "ingest.cuh"
#include "pomfret_gpu.h"
#include "htslib/kfunc.h"

using namespace pomfret_gpu;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "[E::pomfret_gpu] %s failed at %s:%d: %s\n", #call, __FILE__, __LINE__, \
                    cudaGetErrorString(e_));                                                      \
            return POMFRET_GPU_ERR_CUDA;                                                          \
        }                                                                                         \
    } while (0)

namespace {

// Device memory of a batch comes from one arena: a bump allocator over a few large cudaMalloc chunks that
// is rewound by batch_reset().  cudaMalloc / cudaFree serialise across host threads in the driver, so the
// steady state (one chunk, no allocation calls per batch) is what lets several workers share a device.
struct DevArena {
    struct Chunk { uint8_t *p; size_t cap, used; };
    std::vector<Chunk> chunks;
    uint64_t epoch = 1;
    size_t high_water = 0, used_total = 0;
    void *alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (!chunks.empty()) {
            Chunk &c = chunks.back();
            if (c.used + bytes <= c.cap) { void *r = c.p + c.used; c.used += bytes; used_total += bytes; return r; }
        }
        // (large steps: every cudaMalloc / cudaFree is a driver-wide lock, and the rewind below frees with a device-wide sync)
        size_t want = std::max<size_t>(bytes, std::max<size_t>((size_t)256 << 20, chunks.empty() ? 0 : chunks.back().cap));
        void *p = nullptr;
        if (cudaMalloc(&p, want) != cudaSuccess) return nullptr;
        chunks.push_back({(uint8_t *)p, want, bytes});
        used_total += bytes;
        return p;
    }
    void rewind() {
        high_water = std::max(high_water, used_total);
        if (chunks.size() > 1) {  // coalesce: next batch of this size fits one chunk
            for (Chunk &c : chunks) cudaFree(c.p);
            chunks.clear();
            void *p = nullptr;
            size_t want = high_water + high_water / 4;
            if (cudaMalloc(&p, want) == cudaSuccess) chunks.push_back({(uint8_t *)p, want, 0});
        } else if (!chunks.empty()) chunks[0].used = 0;
        used_total = 0;
        epoch++;
    }
    void release() { for (Chunk &c : chunks) cudaFree(c.p); chunks.clear(); }
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    uint64_t epoch = 0;
    DevArena *arena = nullptr;
    // contents survive only while the request fits the current allocation of the current batch epoch
    int ensure(size_t bytes) {
        if (epoch == arena->epoch && bytes <= cap) return 0;
        size_t want = bytes + bytes / 8 + 256;
        p = arena->alloc(want);
        if (!p) { cap = 0; return POMFRET_GPU_ERR_NOMEM; }
        cap = want;
        epoch = arena->epoch;
        return 0;
    }
    void release() { p = nullptr; cap = 0; epoch = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    uint8_t *p = nullptr;
    size_t cap = 0, len = 0;
    size_t min_cap = (size_t)1 << 20;  // cudaMallocHost is slow and serialises: grow in big steps
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        size_t want = std::max(std::max(bytes + bytes / 2, cap * 2), min_cap);
        void *q = nullptr;
        if (cudaMallocHost(&q, want) != cudaSuccess) return POMFRET_GPU_ERR_NOMEM;
        // `len` is the laid-out size; records gathered by the device have no bytes here, so it may exceed cap
        if (p) { memcpy(q, p, len < cap ? len : cap); cudaFreeHost(p); }
        p = (uint8_t *)q; cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = len = 0; }
};

template <typename T> struct PinVec {
    PinBuf b;
    size_t n = 0;
    int push(const T &v) {
        if (int rc = b.reserve((n + 1) * sizeof(T))) return rc;
        reinterpret_cast<T *>(b.p)[n++] = v;
        b.len = n * sizeof(T);
        return 0;
    }
    int resize(size_t m) {
        if (int rc = b.reserve(m * sizeof(T))) return rc;
        n = m; b.len = n * sizeof(T);
        return 0;
    }
    T *data() { return reinterpret_cast<T *>(b.p); }
    T &operator[](size_t i) { return reinterpret_cast<T *>(b.p)[i]; }
    void clear() { n = 0; b.len = 0; }
    void release() { b.release(); n = 0; }
};

// Helper threads for the gather copy of add_reads(): record payloads are scattered over the caller's
// memory (one BAM record each) and have to land in the pinned arena before the DMA engine can take
// them; one core moves ~10 GB/s, PCIe 5 wants ~55 GB/s.  Helpers spin briefly between calls (a loader
// calls add_reads back to back) and then block on a condition variable.
class StagePool {
 public:
    explicit StagePool(int n_helpers) {
        for (int i = 0; i < n_helpers; i++) th_.emplace_back([this] { loop(); });
    }
    ~StagePool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; gen_.fetch_add(1); }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void run(size_t n, const std::function<void(size_t)> &fn) {
        if (th_.empty() || n < 2) { for (size_t i = 0; i < n; i++) fn(i); return; }
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn; n_ = n; next_.store(0); active_.store((int)th_.size());
            gen_.fetch_add(1);
        }
        cv_.notify_all();
        for (size_t i; (i = next_.fetch_add(1)) < n;) fn(i);
        while (active_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    }

 private:
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            // spin, then sleep
            bool got = false;
            for (int spin = 0; spin < 400000 && !got; spin++) got = gen_.load(std::memory_order_acquire) != seen;  // ~0.2 ms
            if (!got) {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_.load() != seen; });
            }
            const std::function<void(size_t)> *fn;
            size_t n;
            {
                std::lock_guard<std::mutex> g(mu_);
                seen = gen_.load();
                if (stop_) return;
                fn = fn_; n = n_;
            }
            for (size_t i; (i = next_.fetch_add(1)) < n;) (*fn)(i);
            active_.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t n_ = 0;
    std::atomic<size_t> next_{0};
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> active_{0};
    bool stop_ = false;
};

enum Stage { ST_EMPTY = 0, ST_SUBMITTED = 1, ST_DECODED = 2, ST_PILED = 3, ST_JOINED = 4, ST_HAPTAGGED = 5 };

}  // namespace

struct HostRegion { uintptr_t begin, end; uint64_t dev; bool ours = true; };  // registered caller memory and its device-visible address (ours: registered here, unregistered here)

struct pomfret_gpu_ctx {
    std::vector<int> devices;
    int n_workers = 1;
    std::mutex mu;
    std::vector<HostRegion> regions;  // sorted by begin
};

struct pomfret_gpu_batch {
    pomfret_gpu_ctx *ctx = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaStream_t side[3] = {};     // join launch groups beside the first one
    cudaEvent_t ev_side[3] = {};
    cudaEvent_t ev[10] = {};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int sm_count = 148;
    int stage = ST_EMPTY;
    // host staging (pinned)
    PinBuf h_blob;
    PinVec<ReadRec> h_reads;
    PinVec<WindowRec> h_win;
    PinVec<uint32_t> h_read_win;
    PinVec<uint32_t> h_win_base, h_win_tile_first;
    PinVec<TileRec> h_tiles;
    PinVec<WindowState> h_state;
    PinVec<uint32_t> h_u32;   // scratch for small D2H reads
    PinVec<uint32_t> h_cta;   // join launch order: window * 2 + direction
    PinVec<uint32_t> h_order_len;  // record indices, longest first
    std::vector<uint32_t> h_end;  // per read: reference end computed on the host (tile planning only)
    // compressed ingest
    PinBuf h_comp;                   // BGZF blocks as they lie in the file
    PinVec<uint32_t> h_ing;          // block and stream tables on their way to the device
    std::vector<pomfret_gpu_bgzf_stream> ing_streams;
    uint32_t ing_records = 0;
    bool ing_timed = false, device_reads = false;
    cudaEvent_t ev_ing[4] = {};
    std::vector<uint32_t> h_dup_of;  // per read: earlier batch read with the same payload (decoded once), or kNoDup
    size_t n_dups = 0;
    uint64_t calls_total = 0;
    uint64_t alg_decode_bytes = 0, alg_haptag_bytes = 0;
    StagePool *pool = nullptr;      // gather-copy helpers (created on first bulk add)
    size_t blob_hint = 0;           // blob size of the previous batch: device buffer is ready before staging starts
    size_t blob_sent = 0;           // bytes of the blob already handed to the DMA engine while staging
    bool streaming = false, h2d_started = false;
    bool direct_any = false, copied_any = false;  // records gathered by the device / copied by the host in this batch
    PinVec<GatherSrc> h_gsrc;
    // device
    DevArena arena;
    DevBuf d_blob, d_reads, d_win, d_read_win, d_calls_pos, d_calls_cat, d_tmp_rank, d_tmp_mpos, d_tmp_mcat;
    DevBuf d_r_ncalls, d_r_status, d_r_end, d_r_id, d_rs_src, d_rs_rev, d_rs_hp;
    DevBuf d_ids[4], d_state, d_tiles, d_win_base, d_win_tile_first, d_tile_out, d_tile_count;
    DevBuf d_site_pos, d_site_start[2], d_site_len[2];
    DevBuf d_mm_xl[2], d_mm_xr[2], d_mm_off[2], d_mm_n[2], d_mm_start[2], d_pool_total, d_mmr_pool, d_ent_pool, d_tab;
    DevBuf d_tags[2], d_order[2];
    DevBuf d_known, d_bases, d_known_first, d_hap_tag, d_hap_status, d_flags, d_cta, d_order_len, d_gsrc, d_dup_of;
    DevBuf d_comp, d_inflated, d_ing_tab, d_ing_small, d_rec_off, d_rec_stream, d_sliced, d_hap_counts, d_ing_cov, d_retag_off, d_retag_val, d_retag_out, d_generic;
    uint32_t pool_cap = 0, tab_sites = 0, site_total = 0, max_sites = 0, max_win_reads = 0;
    pomfret_gpu_config cfg = {};
    uint32_t lo = 0, hi = 0;
    pomfret_gpu_timing tm = {};
    bool have_results = false;
    std::vector<uint8_t> host_tags_fwd, host_tags_bwd;
    std::vector<int32_t> host_rid;
    std::vector<DevBuf *> all_bufs() {
        return {&d_blob, &d_reads, &d_win, &d_read_win, &d_calls_pos, &d_calls_cat, &d_tmp_rank,
                     &d_tmp_mpos, &d_tmp_mcat, &d_r_ncalls, &d_r_status, &d_r_end, &d_r_id, &d_rs_src,
                     &d_rs_rev, &d_rs_hp, &d_ids[0], &d_ids[1], &d_ids[2], &d_ids[3], &d_state,
                     &d_tiles, &d_win_base, &d_win_tile_first, &d_tile_out, &d_tile_count, &d_site_pos,
                     &d_site_start[0], &d_site_start[1], &d_site_len[0], &d_site_len[1], &d_mm_xl[0],
                     &d_mm_xl[1], &d_mm_xr[0], &d_mm_xr[1], &d_mm_off[0], &d_mm_off[1], &d_mm_n[0],
                     &d_mm_n[1], &d_mm_start[0], &d_mm_start[1], &d_pool_total, &d_mmr_pool, &d_ent_pool,
                     &d_tab, &d_tags[0], &d_tags[1], &d_order[0], &d_order[1], &d_known, &d_bases,
                     &d_known_first, &d_hap_tag, &d_hap_status, &d_flags, &d_cta, &d_order_len, &d_gsrc, &d_dup_of,
                     &d_comp, &d_inflated, &d_ing_tab, &d_ing_small, &d_rec_off, &d_rec_stream, &d_sliced, &d_hap_counts, &d_ing_cov, &d_retag_off, &d_retag_val, &d_retag_out, &d_generic};
    }
};

static const int kSideStreams = 3;
static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
static const uint32_t kNoDup = 0xffffffffu;
// dynamic shared memory of join_kernel: one CTA per SM may take kJoinSmemMax, two CTAs per SM kJoinSmemHalf each
static const size_t kJoinSmemMax = (size_t)216 * 1024, kJoinSmemHalf = (size_t)108 * 1024;

extern "C" {

const char *pomfret_gpu_version(void) { return "pomfret_b200 0.1 (sm_100a)"; }

const char *pomfret_gpu_strerror(int rc) {
    switch (rc) {
    case POMFRET_GPU_OK: return "ok";
    case POMFRET_GPU_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case POMFRET_GPU_ERR_CUDA: return "CUDA runtime error";
    case POMFRET_GPU_ERR_NOMEM: return "out of memory";
    case POMFRET_GPU_ERR_ARG: return "invalid argument";
    case POMFRET_GPU_ERR_STATE: return "call out of order";
    case POMFRET_GPU_ERR_FATAL_CIGAR: return "fatal: unknown cigar operation";
    case POMFRET_GPU_ERR_DUP_QNAME: return "duplicated read name";
    case POMFRET_GPU_ERR_UNSUPPORTED: return "input outside the engine's compiled limits";
    case POMFRET_GPU_ERR_MISSING_MD: return "MD tag missing";
    case POMFRET_GPU_ERR_BAD_MD: return "invalid MD";
    default: return "unknown error";
    }
}

int pomfret_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pomfret_gpu_init(pomfret_gpu_ctx **out, const int *devices, int n_devices, int n_workers) {
    if (!out) return POMFRET_GPU_ERR_ARG;
    *out = nullptr;
    int n = pomfret_gpu_device_count();
    if (n <= 0) return POMFRET_GPU_ERR_NO_DEVICE;
    pomfret_gpu_ctx *c = new pomfret_gpu_ctx();
    if (devices && n_devices > 0) {
        for (int i = 0; i < n_devices; i++) {
            if (devices[i] < 0 || devices[i] >= n) { delete c; return POMFRET_GPU_ERR_ARG; }
            c->devices.push_back(devices[i]);
        }
    } else for (int i = 0; i < n; i++) c->devices.push_back(i);
    c->n_workers = n_workers > 0 ? n_workers : 1;
    // create the contexts now (a front end calls init from a start-up thread while it opens its inputs)
    for (int d : c->devices) {
        if (cudaSetDevice(d) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { delete c; return POMFRET_GPU_ERR_CUDA; }
    }
    *out = c;
    return POMFRET_GPU_OK;
}

void pomfret_gpu_destroy(pomfret_gpu_ctx *ctx) { delete ctx; }

int pomfret_gpu_host_register(pomfret_gpu_ctx *ctx, void *ptr, size_t bytes) {
    if (!ctx || !ptr || !bytes) return POMFRET_GPU_ERR_ARG;
    CK(cudaSetDevice(ctx->devices[0]));
    // memory that is pinned already (cudaHostAlloc / cudaMallocHost, e.g. a torch pinned tensor) is taken as it is:
    // it is device-visible under unified addressing and stays the caller's to free
    bool ours = true;
    void *dev = nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        ours = false;
        dev = at.devicePointer;
    } else {
        cudaGetLastError();
        CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
        if (cudaHostGetDevicePointer(&dev, ptr, 0) != cudaSuccess || !dev) { cudaGetLastError(); cudaHostUnregister(ptr); return POMFRET_GPU_ERR_CUDA; }
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    HostRegion r{(uintptr_t)ptr, (uintptr_t)ptr + bytes, (uint64_t)(uintptr_t)dev, ours};
    ctx->regions.insert(std::upper_bound(ctx->regions.begin(), ctx->regions.end(), r,
                                         [](const HostRegion &a, const HostRegion &c) { return a.begin < c.begin; }), r);
    return POMFRET_GPU_OK;
}

int pomfret_gpu_host_unregister(pomfret_gpu_ctx *ctx, void *ptr) {
    if (!ctx || !ptr) return POMFRET_GPU_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        auto it = std::find_if(ctx->regions.begin(), ctx->regions.end(), [&](const HostRegion &r) { return r.begin == (uintptr_t)ptr; });
        if (it == ctx->regions.end()) return POMFRET_GPU_ERR_ARG;
        const bool ours = it->ours;
        ctx->regions.erase(it);
        if (!ours) return POMFRET_GPU_OK;
    }
    CK(cudaHostUnregister(ptr));
    return POMFRET_GPU_OK;
}

// device-visible address of [p, p+n) if it lies inside one registered region, else 0
static uint64_t region_lookup(const std::vector<HostRegion> &regs, const void *p, size_t n, size_t *hint) {
    const uintptr_t a = (uintptr_t)p;
    if (*hint < regs.size() && a >= regs[*hint].begin && a + n <= regs[*hint].end) return regs[*hint].dev + (a - regs[*hint].begin);
    auto it = std::upper_bound(regs.begin(), regs.end(), a, [](uintptr_t v, const HostRegion &r) { return v < r.begin; });
    if (it == regs.begin()) return 0;
    --it;
    if (a + n > it->end) return 0;
    *hint = (size_t)(it - regs.begin());
    return it->dev + (a - it->begin);
}

int pomfret_gpu_batch_begin(pomfret_gpu_ctx *ctx, int worker, int device, pomfret_gpu_batch **out) {
    (void)worker;
    if (!ctx || !out) return POMFRET_GPU_ERR_ARG;
    if (std::find(ctx->devices.begin(), ctx->devices.end(), device) == ctx->devices.end()) return POMFRET_GPU_ERR_ARG;
    CK(cudaSetDevice(device));
    pomfret_gpu_batch *b = new pomfret_gpu_batch();
    b->ctx = ctx;
    b->device = device;
    for (DevBuf *d : b->all_bufs()) d->arena = &b->arena;
    b->h_blob.min_cap = (size_t)64 << 20;
    CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&b->stream2, cudaStreamNonBlocking));
    for (int i = 0; i < kSideStreams; i++) {
        CK(cudaStreamCreateWithFlags(&b->side[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&b->ev_side[i], cudaEventDisableTiming));
    }
    for (auto &e : b->ev) CK(cudaEventCreate(&e));
    for (auto &e : b->ev_ing) CK(cudaEventCreate(&e));
    CK(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
    {
        // (device attributes and kernel attributes are per device, not per batch: looked up once per process;
        //  cudaGetDeviceProperties alone costs tens of milliseconds and serialises the workers that start together)
        static std::mutex mu;
        static std::vector<int> sm_of;
        std::lock_guard<std::mutex> g(mu);
        if ((int)sm_of.size() <= device) sm_of.resize((size_t)device + 1, 0);
        if (!sm_of[(size_t)device]) {
            int n_sm = 0;
#ifndef POMFRET_CUDA_EMU
            if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n_sm = 0;
            CK(cudaFuncSetAttribute(pileup_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PILE_TILE * 4)));
            CK(cudaFuncSetAttribute(join_kernel<true, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJoinSmemMax));
            CK(cudaFuncSetAttribute(join_kernel<true, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJoinSmemMax));
            CK(cudaFuncSetAttribute(join_kernel<false, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJoinSmemMax));
#endif
            sm_of[(size_t)device] = n_sm > 0 ? n_sm : 148;
        }
        b->sm_count = sm_of[(size_t)device];
    }
    *out = b;
    return POMFRET_GPU_OK;
}

int pomfret_gpu_batch_reset(pomfret_gpu_batch *b) {
    if (!b) return POMFRET_GPU_ERR_ARG;
    b->h_blob.len = 0;
    b->h_reads.clear(); b->h_win.clear(); b->h_read_win.clear(); b->h_gsrc.clear();
    b->direct_any = b->copied_any = false;
    b->h_end.clear();
    b->h_dup_of.clear();
    b->n_dups = 0;
    b->ing_records = 0; b->ing_streams.clear(); b->ing_timed = false; b->device_reads = false;
    b->calls_total = 0;
    b->alg_decode_bytes = b->alg_haptag_bytes = 0;
    b->stage = ST_EMPTY;
    b->have_results = false;
    memset(&b->tm, 0, sizeof(b->tm));
    if (cudaSetDevice(b->device) != cudaSuccess) return POMFRET_GPU_ERR_CUDA;
    cudaStreamSynchronize(b->stream);
    b->arena.rewind();
    // A batch of about the size of the previous one gets its device blob up front, so that add_reads()
    // can hand finished parts of the pinned arena to the DMA engine while the caller is still staging.
    b->blob_sent = 0;
    b->h2d_started = false;
    b->streaming = b->blob_hint > 0 && b->d_blob.ensure(b->blob_hint + 1024) == 0;
    return POMFRET_GPU_OK;
}

void pomfret_gpu_batch_end(pomfret_gpu_batch *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    std::vector<DevBuf *> all = b->all_bufs();
    for (DevBuf *d : all) d->release();
    b->arena.release();
    b->h_blob.release(); b->h_reads.release(); b->h_win.release(); b->h_read_win.release();
    b->h_win_base.release(); b->h_win_tile_first.release(); b->h_tiles.release(); b->h_state.release(); b->h_u32.release(); b->h_cta.release(); b->h_order_len.release(); b->h_gsrc.release();
    for (auto &e : b->ev) if (e) cudaEventDestroy(e);
    for (auto &e : b->ev_ing) if (e) cudaEventDestroy(e);
    b->h_comp.release(); b->h_ing.release();
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->stream2) cudaStreamDestroy(b->stream2);
    for (int i = 0; i < kSideStreams; i++) { if (b->ev_side[i]) cudaEventDestroy(b->ev_side[i]); if (b->side[i]) cudaStreamDestroy(b->side[i]); }
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b->pool;
    delete b;
}

// Phase 1 of staging (serial, no payload touched): lay the record's fields out in the blob and reserve
// its call slots.  Every field starts on a 16-byte boundary and is zero padded to the next one.
static int plan_read(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r, size_t *len, int64_t same_as) {
    if (same_as >= 0) {
        // The same alignment record as an earlier read of this batch (it lies in two windows): the slot shares
        // the payload, the call slots and the decode result of that read; only the haplotag is its own.
        if ((uint64_t)same_as >= b->h_reads.n) return POMFRET_GPU_ERR_ARG;
        uint32_t root = (uint32_t)same_as;
        if (b->h_dup_of[root] != kNoDup) root = b->h_dup_of[root];
        ReadRec R = b->h_reads[root];
        if (R.pos != r->pos || R.l_qseq != r->l_qseq || R.n_cigar != r->n_cigar) return POMFRET_GPU_ERR_ARG;
        R.hp = r->hp;
        int rc;
        if ((rc = b->h_reads.push(R))) return rc;
        if ((rc = b->h_read_win.push(0xffffffffu))) return rc;
        b->h_dup_of.push_back(root);
        b->n_dups++;
        return POMFRET_GPU_OK;
    }
    if (r->l_qseq >= (1u << 28)) return POMFRET_GPU_ERR_UNSUPPORTED;  // decode.cuh: DEC_SAT
    ReadRec R;
    memset(&R, 0, sizeof(R));
    R.pos = r->pos; R.l_qseq = r->l_qseq; R.n_cigar = r->n_cigar;
    R.flags = r->flag;
    if (r->tags_malformed) R.flags |= RF_MALFORMED;
    R.hp = r->hp; R.mn = r->mn;
    auto place = [&](size_t n, uint32_t *off16) { *off16 = (uint32_t)(*len / 16); *len = align16(*len + n); };
    place((size_t)r->n_cigar * 4, &R.cigar_off);
    place(((size_t)r->l_qseq + 1) / 2, &R.seq_off);
    if (r->mm) { R.flags |= RF_HAS_MM; R.mm_len = r->mm_len; place(r->mm_len, &R.mm_off); }
    if (r->ml_len >= 0) { R.flags |= RF_HAS_ML; R.ml_len = (uint32_t)r->ml_len; place((size_t)r->ml_len, &R.ml_off); }
    if (r->md) { R.flags |= RF_HAS_MD; R.md_len = r->md_len; place(r->md_len, &R.md_off); }
    if (*len / 16 > 0xfffffff0ull) return POMFRET_GPU_ERR_UNSUPPORTED;  // 32-bit offsets in units of 16 bytes
    // call slots: one per listed base is enough unless implicit canonical calls appear (then the engine
    // re-runs the record with the exact count)
    uint32_t cap = r->ml_len >= 0 ? (uint32_t)r->ml_len : r->mm_len / 2 + 1;
    if (cap < 4) cap = 4;
    R.calls_off = (uint32_t)b->calls_total;
    R.calls_cap = cap;
    b->calls_total += cap;
    if (b->calls_total > 0xfff00000ull) return POMFRET_GPU_ERR_UNSUPPORTED;
    // algorithmic byte counts
    b->alg_decode_bytes += ((uint64_t)r->l_qseq + 1) / 2 + 4ull * r->n_cigar + r->mm_len + (r->ml_len > 0 ? r->ml_len : 0);
    b->alg_haptag_bytes += ((uint64_t)r->l_qseq + 1) / 2 + 4ull * r->n_cigar + r->md_len + 1;
    int rc;
    if ((rc = b->h_reads.push(R))) return rc;
    if ((rc = b->h_read_win.push(0xffffffffu))) return rc;
    b->h_dup_of.push_back(kNoDup);
    return POMFRET_GPU_OK;
}

// Copy into the pinned arena with non-temporal stores: the destination is written once and next read by
// the DMA engine, so it should neither be fetched for ownership nor displace the source from the caches
// (one third less host memory traffic on the staging path, which is what bounds it).  dst is 16-byte aligned.
static inline void copy_stream(uint8_t *dst, const void *src, size_t n) {
#ifdef POMFRET_NT_COPY
    if (n >= 256) {
        const uint8_t *s = (const uint8_t *)src;
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            __m128i a = _mm_loadu_si128((const __m128i *)(s + i)), b = _mm_loadu_si128((const __m128i *)(s + i + 16));
            __m128i c = _mm_loadu_si128((const __m128i *)(s + i + 32)), d = _mm_loadu_si128((const __m128i *)(s + i + 48));
            _mm_stream_si128((__m128i *)(dst + i), a); _mm_stream_si128((__m128i *)(dst + i + 16), b);
            _mm_stream_si128((__m128i *)(dst + i + 32), c); _mm_stream_si128((__m128i *)(dst + i + 48), d);
        }
        if (i < n) memcpy(dst + i, s + i, n - i);
        return;
    }
#endif
    if (n) memcpy(dst, src, n);
}

// Phase 2 (any thread): copy the payload of one planned record into the pinned blob.
static void copy_read(uint8_t *blob, const ReadRec &R, const pomfret_gpu_read_desc *r, uint32_t *end_out) {
    auto put = [&](uint32_t off16, const void *src, size_t n) {
        uint8_t *dst = blob + (size_t)off16 * 16;
        copy_stream(dst, src, n);
        memset(dst + n, 0, align16(n) - n);
    };
    put(R.cigar_off, r->cigar, (size_t)r->n_cigar * 4);
    put(R.seq_off, r->seq, ((size_t)r->l_qseq + 1) / 2);
    // the unused low nibble of an odd-length SEQ must read as "no base" (the kernels do not mask the tail)
    if (r->l_qseq & 1u) blob[(size_t)R.seq_off * 16 + r->l_qseq / 2] &= 0xf0u;
    if (R.flags & RF_HAS_MM) put(R.mm_off, r->mm, r->mm_len);
    if (R.flags & RF_HAS_ML) put(R.ml_off, r->ml, (size_t)r->ml_len);
    if (R.flags & RF_HAS_MD) put(R.md_off, r->md, r->md_len);
    // reference end (tile planning on the host)
    uint64_t rlen = 0;
    for (uint32_t i = 0; i < r->n_cigar; i++) {
        uint32_t op = r->cigar[i] & 15u;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += r->cigar[i] >> 4;
    }
    if (rlen == 0) rlen = 1;
    *end_out = (uint32_t)(r->pos + rlen);
#ifdef POMFRET_NT_COPY
    _mm_sfence();  // the streamed payload is globally visible before the copy engine is pointed at it
#endif
}

// Hand the finished part of the blob to the copy engine (only while the device buffer of this batch is
// known to be large enough; submit() sends whatever is left).
static int stream_blob(pomfret_gpu_batch *b, bool force) {
    if (!b->streaming) return 0;
    if (b->direct_any) { b->streaming = false; return 0; }  // mixed batch: submit() copies the host-staged part in one go
    if (b->h_blob.len + 2048 > b->d_blob.cap) { b->streaming = false; return 0; }  // larger than planned: submit() copies all
    const size_t ready = b->h_blob.len;
    if (ready <= b->blob_sent || (!force && ready - b->blob_sent < ((size_t)8 << 20))) return 0;
    if (cudaSetDevice(b->device) != cudaSuccess) return POMFRET_GPU_ERR_CUDA;
    if (!b->h2d_started) { CK(cudaEventRecord(b->ev[0], b->stream)); b->h2d_started = true; }
    CK(cudaMemcpyAsync(b->d_blob.as<uint8_t>() + b->blob_sent, b->h_blob.p + b->blob_sent, ready - b->blob_sent,
                       cudaMemcpyHostToDevice, b->stream));
    b->tm.bytes_h2d += ready - b->blob_sent;
    b->blob_sent = ready;
    return 0;
}

int pomfret_gpu_batch_add_reads(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r, uint32_t n) {
    return pomfret_gpu_batch_add_reads_shared(b, r, n, nullptr);
}

static int add_reads_impl(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r, uint32_t n, const int64_t *same_as, bool device_ptrs);

int pomfret_gpu_batch_add_reads_shared(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r, uint32_t n, const int64_t *same_as) {
    return add_reads_impl(b, r, n, same_as, false);
}

}  // extern "C"

// device_ptrs: the records come from the compressed ingest; the pointers are device addresses inside the inflated
// streams of this batch (the host must not touch them) and r[i].reserved holds the record's reference end
static int add_reads_impl(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r, uint32_t n, const int64_t *same_as, bool device_ptrs) {
    if (!b || (n && !r)) return POMFRET_GPU_ERR_ARG;
    if (device_ptrs && (b->copied_any || (b->direct_any && !b->device_reads))) return POMFRET_GPU_ERR_STATE;  // one kind of source per batch
    if (!device_ptrs && b->device_reads) return POMFRET_GPU_ERR_STATE;
    if (b->stage != ST_EMPTY) return POMFRET_GPU_ERR_STATE;
    if (n == 0) return POMFRET_GPU_OK;
    const size_t first = b->h_reads.n;
    size_t len = b->h_blob.len;
    const uint64_t calls0 = b->calls_total, dec0 = b->alg_decode_bytes, hap0 = b->alg_haptag_bytes;
    const size_t dups0 = b->n_dups;
    auto rollback = [&]() {  // a failed call leaves the batch as it found it
        b->h_reads.n = first; b->h_reads.b.len = first * sizeof(ReadRec);
        b->h_read_win.n = first; b->h_read_win.b.len = first * 4;
        b->h_dup_of.resize(first);
        b->n_dups = dups0;
        b->calls_total = calls0; b->alg_decode_bytes = dec0; b->alg_haptag_bytes = hap0;
    };
    for (uint32_t i = 0; i < n; i++) {
        if (int rc = plan_read(b, r + i, &len, same_as ? same_as[i] : -1)) { rollback(); return rc; }
    }
    auto is_dup = [&](size_t i) { return b->h_dup_of[first + i] != kNoDup; };
    if (int rc = b->h_gsrc.resize(first + n)) { rollback(); return rc; }
    b->h_end.resize(first + n);
    // records that lie completely inside registered caller buffers stay where they are: the device gathers them
    {
        std::vector<HostRegion> regs;
        if (!device_ptrs) { std::lock_guard<std::mutex> g(b->ctx->mu); regs = b->ctx->regions; }
        bool direct = device_ptrs || !regs.empty();
        size_t hint = 0;
        for (uint32_t i = 0; i < n && direct; i++) {
            const pomfret_gpu_read_desc &d = r[i];
            GatherSrc &G = b->h_gsrc[first + i];
            memset(&G, 0, sizeof(G));
            if (is_dup(i)) continue;
            const void *p[5] = {d.cigar, d.seq, d.mm, d.ml_len >= 0 ? d.ml : nullptr, d.md};
            const size_t sz[5] = {(size_t)d.n_cigar * 4, ((size_t)d.l_qseq + 1) / 2, d.mm ? d.mm_len : 0,
                                  d.ml_len > 0 ? (size_t)d.ml_len : 0, d.md ? d.md_len : 0};
            for (int f = 0; f < 5; f++) {
                if (!p[f] || !sz[f]) continue;
                // (word-aligned reads may touch up to 3 bytes on either side: registered memory is pinned page-wise)
                G.ptr[f] = device_ptrs ? (uint64_t)(uintptr_t)p[f] : region_lookup(regs, p[f], sz[f], &hint);
                if (!G.ptr[f]) { direct = false; break; }
            }
        }
        if (direct) {
            // reference ends (tile planning on the host): the only payload the CPU looks at, a few hundred bytes per record
            uint32_t *ends = b->h_end.data() + first;
            auto scan = [&](size_t i) {
                if (is_dup(i)) return;
                if (device_ptrs) { ends[i] = r[i].reserved; return; }
                uint64_t rlen = 0;
                for (uint32_t c = 0; c < r[i].n_cigar; c++) {
                    const uint32_t op = r[i].cigar[c] & 15u;
                    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += r[i].cigar[c] >> 4;
                }
                ends[i] = (uint32_t)(r[i].pos + (rlen ? rlen : 1));
            };
            if (n >= 64 && !b->pool) {
                unsigned hw = std::thread::hardware_concurrency();
                int per = (int)(hw ? hw : 1) / std::max(1, b->ctx->n_workers);
                if (const char *e = getenv("POMFRET_GPU_STAGE_THREADS")) per = atoi(e);
                b->pool = new StagePool(std::max(0, std::min(per, 12) - 1));
            }
            if (b->pool && n >= 64) b->pool->run((n + 63) / 64, [&](size_t g) { for (size_t i = g * 64; i < std::min<size_t>(n, g * 64 + 64); i++) scan(i); });
            else for (uint32_t i = 0; i < n; i++) scan(i);
            if (b->n_dups != dups0) for (uint32_t i = 0; i < n; i++) if (is_dup(i)) ends[i] = b->h_end[b->h_dup_of[first + i]];
            b->h_blob.len = len;  // layout only: nothing is written on the host
            b->direct_any = true;
            if (device_ptrs) b->device_reads = true;
            return POMFRET_GPU_OK;
        }
        for (uint32_t i = 0; i < n; i++) memset(&b->h_gsrc[first + i], 0, sizeof(GatherSrc));
    }
    b->copied_any = true;
    if (len + 2048 > b->h_blob.cap) {
        // growing the pinned arena moves it: no DMA may still be reading the old one
        if (b->blob_sent) { CK(cudaSetDevice(b->device)); CK(cudaStreamSynchronize(b->stream)); }
        if (int rc = b->h_blob.reserve(len + 2048)) { rollback(); return rc; }
    }
    uint8_t *blob = b->h_blob.p;
    const ReadRec *recs = b->h_reads.data() + first;
    uint32_t *ends = b->h_end.data() + first;
    if (n >= 16 && !b->pool) {
        unsigned hw = std::thread::hardware_concurrency();
        int per = (int)(hw ? hw : 1) / std::max(1, b->ctx->n_workers);
        if (const char *e = getenv("POMFRET_GPU_STAGE_THREADS")) per = atoi(e);
        b->pool = new StagePool(std::max(0, std::min(per, 12) - 1));
    }
    // Groups of ~4 MB: the helpers copy one group while the DMA engine already moves the previous ones.
    // (shared records own no bytes: `laid[i]` is where the first record at or behind i that does begins)
    std::vector<size_t> laid;
    if (b->n_dups != dups0) {
        laid.resize(n);
        size_t cur = len;
        for (uint32_t i = n; i-- > 0;) { if (!is_dup(i)) cur = (size_t)recs[i].cigar_off * 16; laid[i] = cur; }
    }
    auto start_of = [&](uint32_t i) { return laid.empty() ? (size_t)recs[i].cigar_off * 16 : laid[i]; };
    auto copy_one = [&](size_t i) { if (!is_dup(i)) copy_read(blob, recs[i], r + i, ends + i); };
    for (uint32_t g0 = 0; g0 < n;) {
        uint32_t g1 = g0;
        const size_t from = start_of(g0);
        while (g1 < n && (g1 - g0 < 16 || start_of(g1) - from < ((size_t)4 << 20))) g1++;
        if (b->pool && g1 - g0 >= 16) b->pool->run(g1 - g0, [&](size_t i) { copy_one(g0 + i); });
        else for (uint32_t i = g0; i < g1; i++) copy_one(i);
        b->h_blob.len = g1 < n ? start_of(g1) : len;
        if (int rc = stream_blob(b, false)) return rc;
        g0 = g1;
    }
    if (b->n_dups != dups0) for (uint32_t i = 0; i < n; i++) if (is_dup(i)) ends[i] = b->h_end[b->h_dup_of[first + i]];
    return POMFRET_GPU_OK;
}

extern "C" {

int pomfret_gpu_batch_add_read(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r) {
    if (!r) return POMFRET_GPU_ERR_ARG;
    return pomfret_gpu_batch_add_reads(b, r, 1);
}

int pomfret_gpu_batch_add_window(pomfret_gpu_batch *b, uint32_t ref_start, uint32_t ref_end, uint32_t first_read,
                                 uint32_t n_reads) {
    if (!b) return POMFRET_GPU_ERR_ARG;
    if (b->stage != ST_EMPTY) return POMFRET_GPU_ERR_STATE;
    if ((uint64_t)first_read + n_reads > b->h_reads.n) return POMFRET_GPU_ERR_ARG;
    if (n_reads >= 60000) return POMFRET_GPU_ERR_UNSUPPORTED;  // 16-bit count fields
    WindowRec W;
    memset(&W, 0, sizeof(W));
    W.ref_start = ref_start; W.ref_end = ref_end; W.first_read = first_read; W.n_reads = n_reads;
    for (uint32_t i = 0; i < n_reads; i++) {
        if (b->h_read_win[first_read + i] != 0xffffffffu) return POMFRET_GPU_ERR_ARG;  // windows must not share records
        b->h_read_win[first_read + i] = (uint32_t)b->h_win.n;
    }
    return b->h_win.push(W);
}

int pomfret_gpu_batch_add_windows(pomfret_gpu_batch *b, const uint32_t *ref_start, const uint32_t *ref_end,
                                  const uint32_t *first_read, const uint32_t *n_reads, uint32_t n) {
    if (!b || (n && (!ref_start || !ref_end || !first_read || !n_reads))) return POMFRET_GPU_ERR_ARG;
    for (uint32_t i = 0; i < n; i++)
        if (int rc = pomfret_gpu_batch_add_window(b, ref_start[i], ref_end[i], first_read[i], n_reads[i])) return rc;
    return POMFRET_GPU_OK;
}

static int up(pomfret_gpu_batch *b, DevBuf &d, const void *src, size_t bytes) {
    if (int rc = d.ensure(bytes ? bytes : 16)) return rc;
    if (bytes) {
        CK(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, b->stream));
        b->tm.bytes_h2d += bytes;
    }
    return 0;
}

int pomfret_gpu_batch_submit(pomfret_gpu_batch *b) {
    if (!b) return POMFRET_GPU_ERR_ARG;
    if (b->stage != ST_EMPTY) return POMFRET_GPU_ERR_STATE;
    CK(cudaSetDevice(b->device));
    const size_t nr = b->h_reads.n, nw = b->h_win.n;
    int rc;
    // the blob keeps 1 KB of zeroed slack behind the last field
    if (b->copied_any) {
        if (b->h_blob.len + 2048 > b->h_blob.cap) {
            if (b->blob_sent) CK(cudaStreamSynchronize(b->stream));
            if ((rc = b->h_blob.reserve(b->h_blob.len + 2048))) return rc;
        }
        memset(b->h_blob.p + b->h_blob.len, 0, 1024);
        b->h_blob.len += 1024;
        if ((rc = stream_blob(b, true))) return rc;
        if (!b->h2d_started) { CK(cudaEventRecord(b->ev[0], b->stream)); b->h2d_started = true; }
        if (!b->streaming) {
            b->blob_sent = 0;
            if ((rc = up(b, b->d_blob, b->h_blob.p, b->h_blob.len))) return rc;
        }
    } else {
        // every record is gathered by the device: only the slack has to be prepared
        if (!b->h2d_started) { CK(cudaEventRecord(b->ev[0], b->stream)); b->h2d_started = true; }
        if ((rc = b->d_blob.ensure(b->h_blob.len + 2048))) return rc;
        CK(cudaMemsetAsync(b->d_blob.as<uint8_t>() + b->h_blob.len, 0, 1024, b->stream));
        b->h_blob.len += 1024;
    }
    b->blob_hint = std::max(b->h_blob.len, b->blob_hint - b->blob_hint / 16);  // follows the batch size, decays slowly
    b->h_blob.len -= 1024;
    if ((rc = up(b, b->d_reads, b->h_reads.data(), nr * sizeof(ReadRec)))) return rc;
    if (b->direct_any && nr) {
        if ((rc = up(b, b->d_gsrc, b->h_gsrc.data(), nr * sizeof(GatherSrc)))) return rc;
        GatherParams G;
        G.reads = b->d_reads.as<ReadRec>(); G.src = b->d_gsrc.as<GatherSrc>(); G.blob = b->d_blob.as<uint8_t>(); G.n_reads = (uint32_t)nr;
        POMFRET_LAUNCH(gather_kernel, (unsigned)((nr + GATHER_WARPS - 1) / GATHER_WARPS), GATHER_WARPS * 32, 0, b->stream, G);
        b->tm.launches++;
        for (size_t i = 0; i < nr && !b->device_reads; i++)  // (ingested records crossed the bus compressed: counted there)
            for (int f = 0; f < 5; f++)
                if (b->h_gsrc[i].ptr[f]) {
                    const ReadRec &R = b->h_reads[i];
                    const uint64_t sz[5] = {(uint64_t)R.n_cigar * 4, ((uint64_t)R.l_qseq + 1) / 2, R.mm_len, R.ml_len, R.md_len};
                    b->tm.bytes_h2d += sz[f];
                }
        CK(cudaGetLastError());
    }
    if ((rc = up(b, b->d_win, b->h_win.data(), nw * sizeof(WindowRec)))) return rc;
    if ((rc = up(b, b->d_read_win, b->h_read_win.data(), nr * 4))) return rc;
    // queue order of the per-record kernels: longest records first (counting sort over 256-base buckets)
    // (records shared with an earlier read come last: decode stops in front of them, the per-slot kernels go on)
    if ((rc = b->h_order_len.resize(nr ? nr : 1))) return rc;
    {
        constexpr uint32_t NB = 4096;
        std::vector<uint32_t> cnt(2 * NB + 1, 0);
        auto bucket = [&](size_t i) { return (b->h_dup_of[i] != kNoDup ? NB : 0u) + NB - 1 - std::min<uint32_t>(b->h_reads[i].l_qseq >> 8, NB - 1); };
        for (size_t i = 0; i < nr; i++) cnt[bucket(i)]++;
        uint32_t acc = 0;
        for (uint32_t k = 0; k <= 2 * NB; k++) { uint32_t c = cnt[k]; cnt[k] = acc; acc += c; }
        for (size_t i = 0; i < nr; i++) b->h_order_len[cnt[bucket(i)]++] = (uint32_t)i;
    }
    if ((rc = up(b, b->d_order_len, b->h_order_len.data(), nr * 4))) return rc;
    if (b->n_dups && (rc = up(b, b->d_dup_of, b->h_dup_of.data(), nr * 4))) return rc;
    CK(cudaEventRecord(b->ev[1], b->stream));
    b->stage = ST_SUBMITTED;
    return POMFRET_GPU_OK;
}

int pomfret_gpu_batch_rewind(pomfret_gpu_batch *b) {
    if (!b) return POMFRET_GPU_ERR_ARG;
    if (b->stage < ST_SUBMITTED) return POMFRET_GPU_ERR_STATE;
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    const uint64_t h2d = b->tm.bytes_h2d;
    memset(&b->tm, 0, sizeof(b->tm));
    b->tm.bytes_h2d = h2d;
    b->have_results = false;
    b->stage = ST_SUBMITTED;
    return POMFRET_GPU_OK;
}

static int launch_decode(pomfret_gpu_batch *b) {
    const size_t nr = b->h_reads.n;
    const size_t slots = (size_t)b->calls_total + 16;
    int rc;
    if ((rc = b->d_calls_pos.ensure(slots * 4)) || (rc = b->d_calls_cat.ensure(slots)) || (rc = b->d_tmp_rank.ensure(slots * 4)) ||
        (rc = b->d_tmp_mpos.ensure(slots * 4)) || (rc = b->d_tmp_mcat.ensure(slots)) || (rc = b->d_r_ncalls.ensure(nr * 4 + 16)) ||
        (rc = b->d_r_status.ensure(nr * 4 + 16)) || (rc = b->d_r_end.ensure(nr * 4 + 16)) || (rc = b->d_flags.ensure(16)))
        return rc;
    CK(cudaMemsetAsync(b->d_flags.p, 0, 16, b->stream));
    DecodeParams P;
    P.reads = b->d_reads.as<ReadRec>();
    P.n_reads = (uint32_t)nr;
    P.blob = b->d_blob.as<uint8_t>();
    P.calls_pos = b->d_calls_pos.as<uint32_t>();
    P.calls_cat = b->d_calls_cat.as<uint8_t>();
    P.tmp_rank = b->d_tmp_rank.as<uint32_t>();
    P.tmp_mpos = b->d_tmp_mpos.as<uint32_t>();
    P.tmp_mcat = b->d_tmp_mcat.as<uint8_t>();
    P.r_ncalls = b->d_r_ncalls.as<uint32_t>();
    P.r_status = b->d_r_status.as<uint32_t>();
    P.r_end = b->d_r_end.as<uint32_t>();
    P.n_overflow = b->d_flags.as<uint32_t>();
    P.next = b->d_flags.as<uint32_t>() + 1;
    P.order = nr ? b->d_order_len.as<uint32_t>() : nullptr;
    P.lo = b->lo; P.hi = b->hi;
    P.no_lean = 0;
    if (const char *e = getenv("POMFRET_GPU_DECODE_LEAN")) P.no_lean = !strcmp(e, "0");  // test hook: streaming path for every record
    // records for the general sequential path are collected and run one per thread by decode_generic_kernel
    // (POMFRET_GPU_DECODE_GENERIC=inplace: lane 0 of the record's warp runs them, the round-1 form; measurement hook)
    P.generic_list = nullptr;
    P.n_generic = b->d_flags.as<uint32_t>() + 2;
    {
        const char *e = getenv("POMFRET_GPU_DECODE_GENERIC");
        if (!(e && !strcmp(e, "inplace")) && nr) {
            if ((rc = b->d_generic.ensure(nr * 4 + 16))) return rc;
            P.generic_list = b->d_generic.as<uint32_t>();
        }
    }
    const size_t nq = nr - b->n_dups;  // the queue holds every distinct record once
    P.n_queue = (uint32_t)nq;
    if (nq) {
        // one CTA per 4 records, at most the resident set (8 CTAs per SM): the warps loop over a queue
        unsigned grid = (unsigned)std::min<size_t>((nq + DEC_WARPS - 1) / DEC_WARPS, (size_t)b->sm_count * 8);
        if (const char *e = getenv("POMFRET_GPU_DECODE_QUEUE")) {  // measurement hook: "0" = one warp per record in batch order
            if (!strcmp(e, "0")) { grid = (unsigned)((nr + DEC_WARPS - 1) / DEC_WARPS); P.order = nullptr; P.n_queue = (uint32_t)nr; }
        }
        POMFRET_LAUNCH(decode_kernel, grid, DEC_WARPS * 32, 0, b->stream, P);
        b->tm.launches++;
        if (P.generic_list) {  // (the count stays on the device: threads beyond it leave at once)
            POMFRET_LAUNCH(decode_generic_kernel, (unsigned)((nq + GEN_THREADS - 1) / GEN_THREADS), GEN_THREADS, 0, b->stream, P);
            b->tm.launches++;
        }
    }
    if (b->n_dups) {
        POMFRET_LAUNCH(share_decoded_kernel, (unsigned)((nr + 255) / 256), 256, 0, b->stream, b->d_dup_of.as<uint32_t>(), (uint32_t)nr,
                       P.r_ncalls, P.r_status, P.r_end);
        b->tm.launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

int pomfret_gpu_decode(pomfret_gpu_batch *b, uint8_t qual_lo, uint8_t qual_hi) {
    if (!b) return POMFRET_GPU_ERR_ARG;
    if (b->stage != ST_SUBMITTED) return POMFRET_GPU_ERR_STATE;
    CK(cudaSetDevice(b->device));
    b->lo = qual_lo; b->hi = qual_hi;
    CK(cudaEventRecord(b->ev[2], b->stream));
    if (int rc = launch_decode(b)) return rc;
    CK(cudaEventRecord(b->ev[3], b->stream));
    b->stage = ST_DECODED;
    return POMFRET_GPU_OK;
}

// read sets, pileup tiles, site layout, methmer sizing — everything up to the pool-size round trip
static int launch_pileup_stages(pomfret_gpu_batch *b) {
    const size_t nr = b->h_reads.n, nw = b->h_win.n;
    const pomfret_gpu_config &cfg = b->cfg;
    int rc;
    // ---- plan: site capacity per window, position tiles ----
    b->h_win_base.clear(); b->h_win_tile_first.clear(); b->h_tiles.clear();
    uint64_t site_total = 0, tile_out_total = 0;
    const uint32_t cov = (uint32_t)std::max(cfg.cov_for_selection, 1);
    for (size_t w = 0; w < nw; w++) {
        WindowRec &W = b->h_win[w];
        uint64_t caps = 0;
        uint32_t mn = 0xffffffffu, mx = 0;
        for (uint32_t i = 0; i < W.n_reads; i++) {
            const ReadRec &R = b->h_reads[W.first_read + i];
            caps += R.calls_cap;
            mn = std::min(mn, R.pos);
            mx = std::max(mx, b->h_end[W.first_read + i]);
        }
        W.site_off = (uint32_t)site_total;
        W.site_cap = (uint32_t)(caps / (2ull * cov) + 2);
        site_total += W.site_cap;
        uint32_t base = W.n_reads ? mn - 1u : 0u;  // calls lie in [read start - 1, read end]
        uint32_t range = W.n_reads ? mx - base + 2u : 0u;
        uint32_t n_tiles = (range + PILE_TILE - 1) / PILE_TILE;
        if ((rc = b->h_win_base.push(base)) || (rc = b->h_win_tile_first.push((uint32_t)b->h_tiles.n))) return rc;
        for (uint32_t t = 0; t < n_tiles; t++) {
            TileRec T;
            T.window = (uint32_t)w; T.tile = t;
            T.out_off = (uint32_t)tile_out_total;
            T.out_cap = std::min<uint32_t>(W.site_cap, PILE_TILE);
            tile_out_total += T.out_cap;
            if ((rc = b->h_tiles.push(T))) return rc;
        }
        if (site_total > 0xfff00000ull || tile_out_total > 0xfff00000ull) return POMFRET_GPU_ERR_UNSUPPORTED;
    }
    if ((rc = b->h_win_tile_first.push((uint32_t)b->h_tiles.n))) return rc;
    b->site_total = (uint32_t)site_total;
    if ((rc = up(b, b->d_win, b->h_win.data(), nw * sizeof(WindowRec)))) return rc;
    if ((rc = up(b, b->d_win_base, b->h_win_base.data(), nw * 4))) return rc;
    if ((rc = up(b, b->d_win_tile_first, b->h_win_tile_first.data(), (nw + 1) * 4))) return rc;
    if ((rc = up(b, b->d_tiles, b->h_tiles.data(), b->h_tiles.n * sizeof(TileRec)))) return rc;
    const size_t n4 = nr * 4 + 16;
    if ((rc = b->d_r_id.ensure(n4)) || (rc = b->d_rs_src.ensure(n4)) || (rc = b->d_rs_rev.ensure(n4)) || (rc = b->d_rs_hp.ensure(n4)) ||
        (rc = b->d_ids[0].ensure(n4)) || (rc = b->d_ids[1].ensure(n4)) || (rc = b->d_ids[2].ensure(n4)) || (rc = b->d_ids[3].ensure(n4)) ||
        (rc = b->d_state.ensure(nw * sizeof(WindowState) + 16)) || (rc = b->d_tile_out.ensure(tile_out_total * 4 + 16)) ||
        (rc = b->d_tile_count.ensure(b->h_tiles.n * 4 + 16)) || (rc = b->d_site_pos.ensure(site_total * 4 + 16)) ||
        (rc = b->d_pool_total.ensure(16)))
        return rc;
    for (int d = 0; d < 2; d++) {
        if ((rc = b->d_site_start[d].ensure(site_total * 4 + 16)) || (rc = b->d_site_len[d].ensure(site_total + 16)) ||
            (rc = b->d_mm_xl[d].ensure(n4)) || (rc = b->d_mm_xr[d].ensure(n4)) || (rc = b->d_mm_off[d].ensure(n4)) ||
            (rc = b->d_mm_n[d].ensure(n4)) || (rc = b->d_mm_start[d].ensure(n4)) || (rc = b->d_tags[d].ensure(nr + 16)) ||
            (rc = b->d_order[d].ensure(n4)))
            return rc;
    }
    CK(cudaMemsetAsync(b->d_pool_total.p, 0, 16, b->stream));
    if (nw == 0) return 0;
    // ---- read sets ----
    ReadsetParams R;
    R.reads = b->d_reads.as<ReadRec>(); R.win = b->d_win.as<WindowRec>(); R.state = b->d_state.as<WindowState>();
    R.r_status = b->d_r_status.as<uint32_t>(); R.r_end = b->d_r_end.as<uint32_t>(); R.r_ncalls = b->d_r_ncalls.as<uint32_t>();
    R.r_id = b->d_r_id.as<int32_t>(); R.rs_src = b->d_rs_src.as<uint32_t>(); R.rs_rev = b->d_rs_rev.as<uint32_t>();
    R.ids_left = b->d_ids[0].as<uint32_t>(); R.ids_left_strict = b->d_ids[1].as<uint32_t>();
    R.ids_right = b->d_ids[2].as<uint32_t>(); R.ids_right_strict = b->d_ids[3].as<uint32_t>();
    R.rs_hp = b->d_rs_hp.as<int32_t>(); R.n_windows = (uint32_t)nw;
    CK(cudaEventRecord(b->ev[4], b->stream));
    POMFRET_LAUNCH(readset_kernel, (unsigned)nw, RS_THREADS, 0, b->stream, R);
    b->tm.launches++;
    CK(cudaEventRecord(b->ev[5], b->stream));
    // ---- pileup tiles + site layout ----
    PileupParams Q;
    Q.win = b->d_win.as<WindowRec>(); Q.state = b->d_state.as<WindowState>(); Q.tiles = b->d_tiles.as<TileRec>();
    Q.win_base = b->d_win_base.as<uint32_t>(); Q.reads = b->d_reads.as<ReadRec>(); Q.rs_src = b->d_rs_src.as<uint32_t>();
    Q.r_ncalls = b->d_r_ncalls.as<uint32_t>(); Q.r_status = b->d_r_status.as<uint32_t>(); Q.r_end = b->d_r_end.as<uint32_t>();
    Q.calls_pos = b->d_calls_pos.as<uint32_t>(); Q.calls_cat = b->d_calls_cat.as<uint8_t>();
    Q.tile_out = b->d_tile_out.as<uint32_t>(); Q.tile_count = b->d_tile_count.as<uint32_t>();
    Q.cov = (uint32_t)cfg.cov_for_selection;
    if (b->h_tiles.n) {
        POMFRET_LAUNCH(pileup_tile_kernel, (unsigned)b->h_tiles.n, PILE_THREADS, PILE_TILE * 4, b->stream, Q);
        b->tm.launches++;
    }
    SitesParams S;
    S.win = b->d_win.as<WindowRec>(); S.state = b->d_state.as<WindowState>(); S.tiles = b->d_tiles.as<TileRec>();
    S.win_tile_first = b->d_win_tile_first.as<uint32_t>(); S.tile_out = b->d_tile_out.as<uint32_t>();
    S.tile_count = b->d_tile_count.as<uint32_t>(); S.site_pos = b->d_site_pos.as<uint32_t>();
    for (int d = 0; d < 2; d++) { S.site_start[d] = b->d_site_start[d].as<uint32_t>(); S.site_len[d] = b->d_site_len[d].as<uint8_t>(); }
    S.k = cfg.k; S.k_span = cfg.k_span;
    POMFRET_LAUNCH(sites_finalize_kernel, (unsigned)nw, 256, 0, b->stream, S);
    b->tm.launches++;
    CK(cudaEventRecord(b->ev[6], b->stream));
    CK(cudaGetLastError());
    return 0;
}

static void fill_methmer_params(pomfret_gpu_batch *b, MethmerParams &M) {
    M.win = b->d_win.as<WindowRec>(); M.state = b->d_state.as<WindowState>(); M.read_win = b->d_read_win.as<uint32_t>();
    M.reads = b->d_reads.as<ReadRec>(); M.rs_src = b->d_rs_src.as<uint32_t>(); M.r_ncalls = b->d_r_ncalls.as<uint32_t>();
    M.r_status = b->d_r_status.as<uint32_t>(); M.calls_pos = b->d_calls_pos.as<uint32_t>(); M.calls_cat = b->d_calls_cat.as<uint8_t>();
    for (int d = 0; d < 2; d++) {
        M.site_start[d] = b->d_site_start[d].as<uint32_t>(); M.site_len[d] = b->d_site_len[d].as<uint8_t>();
        M.mm_xl[d] = b->d_mm_xl[d].as<uint32_t>(); M.mm_xr[d] = b->d_mm_xr[d].as<uint32_t>(); M.mm_off[d] = b->d_mm_off[d].as<uint32_t>();
        M.mm_n[d] = b->d_mm_n[d].as<uint32_t>(); M.mm_start[d] = b->d_mm_start[d].as<uint32_t>();
    }
    M.pool_total = b->d_pool_total.as<uint32_t>(); M.mmr_pool = b->d_mmr_pool.as<uint32_t>(); M.ent_pool = b->d_ent_pool.as<uint32_t>();
    M.pool_cap = b->pool_cap; M.n_slots = (uint32_t)b->h_reads.n; M.k = b->cfg.k;
    M.order = b->d_order_len.as<uint32_t>();
}

int pomfret_gpu_pileup(pomfret_gpu_batch *b, const pomfret_gpu_config *cfg) {
    if (!b || !cfg) return POMFRET_GPU_ERR_ARG;
    if (b->stage != ST_DECODED) return POMFRET_GPU_ERR_STATE;
    if (cfg->k < 1 || cfg->k > kMaxK || cfg->n_candidates_per_iter > JOIN_MAX_CAND || cfg->n_candidates_per_iter < 1)
        return POMFRET_GPU_ERR_UNSUPPORTED;
    CK(cudaSetDevice(b->device));
    b->cfg = *cfg;
    const size_t nr = b->h_reads.n, nw = b->h_win.n;
    int rc;
    if ((rc = launch_pileup_stages(b))) return rc;
    if (nw == 0) { b->stage = ST_PILED; return POMFRET_GPU_OK; }
    MethmerParams M;
    fill_methmer_params(b, M);
    POMFRET_LAUNCH(methmer_size_kernel, (unsigned)nw, RS_THREADS, 0, b->stream, M);
    b->tm.launches++;
    // ---- the one round trip: pool sizes (and whether any record ran out of call slots) ----
    if ((rc = b->h_u32.resize(8)) || (rc = b->h_state.resize(nw))) return rc;
    for (int attempt = 0;; attempt++) {
        CK(cudaMemcpyAsync(b->h_state.data(), b->d_state.p, nw * sizeof(WindowState), cudaMemcpyDeviceToHost, b->stream));
        CK(cudaMemcpyAsync(b->h_u32.data(), b->d_pool_total.p, 12, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaMemcpyAsync(b->h_u32.data() + 4, b->d_flags.p, 4, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
        if (b->h_u32[4] == 0) break;
        if (attempt > 0) return POMFRET_GPU_ERR_UNSUPPORTED;
        // Implicit canonical calls (blockjoin.c:666-700) can exceed one slot per listed base: give the affected
        // records the hard upper bound (every second base + every listed base) and run the stages again.
        std::vector<uint32_t> st(nr);
        CK(cudaMemcpy(st.data(), b->d_r_status.p, nr * 4, cudaMemcpyDeviceToHost));
        uint64_t total = 0;
        for (size_t i = 0; i < nr; i++) {
            ReadRec &R = b->h_reads[i];
            if (b->h_dup_of[i] != kNoDup) {  // shares the (already re-planned) slots of the earlier read
                R.calls_cap = b->h_reads[b->h_dup_of[i]].calls_cap;
                R.calls_off = b->h_reads[b->h_dup_of[i]].calls_off;
                continue;
            }
            if (st[i] & RS_OVERFLOW) R.calls_cap = R.calls_cap + R.l_qseq / 2 + 16;
            R.calls_off = (uint32_t)total;
            total += R.calls_cap;
        }
        if (total > 0xfff00000ull) return POMFRET_GPU_ERR_UNSUPPORTED;
        b->calls_total = total;
        if ((rc = up(b, b->d_reads, b->h_reads.data(), nr * sizeof(ReadRec)))) return rc;
        if ((rc = launch_decode(b))) return rc;
        if ((rc = launch_pileup_stages(b))) return rc;
        fill_methmer_params(b, M);
        POMFRET_LAUNCH(methmer_size_kernel, (unsigned)nw, RS_THREADS, 0, b->stream, M);
        b->tm.launches += 1;
    }
    const uint32_t mmr_total = b->h_u32[0], tab_sites = b->h_u32[1];
    b->pool_cap = mmr_total + 64;
    b->tab_sites = tab_sites;
    b->max_sites = b->h_u32[2];
    const size_t row_words = join_row_stride(cfg->k);
    if ((rc = b->d_mmr_pool.ensure((size_t)b->pool_cap * 4)) || (rc = b->d_ent_pool.ensure((size_t)b->pool_cap * 4)) ||
        (rc = b->d_tab.ensure(((size_t)tab_sites + 1) * row_words * 4)))
        return rc;
    fill_methmer_params(b, M);
    unsigned grid = (unsigned)std::min<size_t>((nr * 2 + MMR_WARPS - 1) / MMR_WARPS, (size_t)b->sm_count * 16);
    if (grid) {
        POMFRET_LAUNCH(methmer_fill_kernel, grid, MMR_WARPS * 32, 0, b->stream, M);
        b->tm.launches++;
    }
    CK(cudaEventRecord(b->ev[7], b->stream));
    CK(cudaGetLastError());
    b->stage = ST_PILED;
    return POMFRET_GPU_OK;
}

int pomfret_gpu_join(pomfret_gpu_batch *b, const pomfret_gpu_config *cfg) {
    if (!b || !cfg) return POMFRET_GPU_ERR_ARG;
    if (b->stage != ST_PILED) return POMFRET_GPU_ERR_STATE;
    if (cfg->k != b->cfg.k) return POMFRET_GPU_ERR_ARG;
    CK(cudaSetDevice(b->device));
    const size_t nw = b->h_win.n;
    JoinParams J;
    J.win = b->d_win.as<WindowRec>(); J.state = b->d_state.as<WindowState>(); J.rs_rev = b->d_rs_rev.as<uint32_t>();
    J.rs_hp = b->d_rs_hp.as<int32_t>();
    J.ids_left = b->d_ids[0].as<uint32_t>(); J.ids_left_strict = b->d_ids[1].as<uint32_t>();
    J.ids_right = b->d_ids[2].as<uint32_t>(); J.ids_right_strict = b->d_ids[3].as<uint32_t>();
    J.site_pos = b->d_site_pos.as<uint32_t>();
    for (int d = 0; d < 2; d++) {
        J.mm_off[d] = b->d_mm_off[d].as<uint32_t>(); J.mm_n[d] = b->d_mm_n[d].as<uint32_t>(); J.mm_start[d] = b->d_mm_start[d].as<uint32_t>();
        J.tags[d] = b->d_tags[d].as<uint8_t>(); J.order[d] = b->d_order[d].as<uint32_t>();
    }
    J.mmr_pool = b->d_mmr_pool.as<uint32_t>(); J.tab = b->d_tab.as<uint32_t>();
    J.n_cand = cfg->n_candidates_per_iter; J.cov_run = cfg->cov_for_runtime; J.k = cfg->k;
    // Shared-memory plan.  Every CTA carries the per-read state of the largest window and the look-ahead key cache;
    // the count tables of a window go to shared memory if they fit: with 8-bit counts (16-bit entries) when fewer
    // than 256 reads touch any site of the window (WindowState::max_cov, the usual case up to ~200x), else with
    // 16-bit counts.  Windows are sorted into launch groups by what a CTA needs, so that small windows run three
    // to an SM and only the largest take an SM for themselves; the groups go to their own streams.  A window
    // whose tables exceed an SM keeps them in the global pool (16-bit counts).
    // one warp per candidate slot (n_cand + 1) plus one that serves the look-ahead slot while the others score
    const unsigned join_threads = 32u * (unsigned)std::min(JOIN_WARPS, std::max(4, J.n_cand + 2));
    const int join_warps = (int)join_threads / 32;
    uint32_t max_reads = 0;
    for (size_t w = 0; w < nw; w++) max_reads = std::max(max_reads, b->h_win[w].n_reads);
    J.meta_cap = max_reads <= 4096 ? max_reads : 0;
    J.stage_cap = join_stage_cap(J.n_cand);
    const uint32_t stride = join_row_stride(cfg->k);
    // (three CTAs of 16 warps fill the register file at the kernel's 40 registers per thread: no tier beyond three per SM)
    constexpr int kTiers = 3;
    size_t tier_limit[kTiers] = {kJoinSmemMax / POMFRET_JOIN_MIN_CTAS - 1024, kJoinSmemHalf, kJoinSmemMax};
    bool allow_u8 = true;
    if (nw * 2 <= (size_t)b->sm_count) tier_limit[0] = tier_limit[1] = kJoinSmemMax;  // every CTA gets an SM anyway
    if (const char *e = getenv("POMFRET_GPU_JOIN_SMEM")) {
        // test hook: "0" keeps per-read state and count tables in global memory (the paths very large windows
        // take), "half" forbids the one-CTA-per-SM launch, "u16" forbids the 8-bit tables
        if (!strcmp(e, "0")) { J.meta_cap = 0; for (size_t &t : tier_limit) t = 0; }
        else if (!strcmp(e, "half")) tier_limit[2] = tier_limit[1];
        else if (!strcmp(e, "u16")) allow_u8 = false;
    }
    // group index: variant (0: 8-bit counts, 1: 16-bit counts) * kTiers + tier; 2 * kTiers: global tables
    constexpr int kGroups = 2 * kTiers + 1;
    std::vector<uint32_t> members[kGroups];
    uint32_t group_entries[kGroups] = {};
    for (size_t w = 0; w < nw; w++) {
        const WindowState &S = b->h_state[w];
        if (S.n == 0 || S.n_sites == 0 || S.status != 0) continue;  // nothing to propagate: no CTA
        const uint32_t entries = S.n_sites * stride;
        int g = 2 * kTiers;
        for (int variant = allow_u8 && S.max_cov < 256u ? 0 : 1; variant < 2 && g == 2 * kTiers; variant++)
            for (int t = 0; t < kTiers; t++)
                if (join_smem_bytes(entries + 1, variant ? 4 : 2, J.meta_cap, J.n_cand, join_warps) <= tier_limit[t]) { g = variant * kTiers + t; break; }
        members[g].push_back((uint32_t)w);
        group_entries[g] = std::max(group_entries[g], entries);
    }
    if (int rc = b->h_cta.resize(nw * 2 + 1)) return rc;
    size_t n_cta = 0, group_first[kGroups];
    for (int g = 0; g < kGroups; g++) {
        group_first[g] = n_cta;
        for (uint32_t w : members[g]) { b->h_cta[n_cta++] = w * 2; b->h_cta[n_cta++] = w * 2 + 1; }
    }
    if (int rc = up(b, b->d_cta, b->h_cta.data(), n_cta * 4)) return rc;
    CK(cudaEventRecord(b->ev[8], b->stream));
    CK(cudaEventRecord(b->ev_fork, b->stream));
    int n_side = 0;
    for (int g = 0; g < kGroups; g++) {
        if (members[g].empty()) continue;
        // the first non-empty group stays on the batch's stream, the others fork to side streams
        cudaStream_t st = b->stream;
        if (n_side > 0) {
            st = b->side[(n_side - 1) % kSideStreams];
            CK(cudaStreamWaitEvent(st, b->ev_fork, 0));
        }
        JoinParams JG = J;
        JG.cta_map = b->d_cta.as<uint32_t>() + group_first[g];
        const unsigned grid = (unsigned)members[g].size() * 2u;
        if (g == 2 * kTiers) {
            JG.smem_tab_words = 0;
            const size_t smem = join_smem_bytes(1, 4, J.meta_cap, J.n_cand, join_warps);
            POMFRET_LAUNCH((join_kernel<false, uint32_t>), grid, join_threads, smem, st, JG);
        } else if (g < kTiers) {
            JG.smem_tab_words = group_entries[g];
            const size_t smem = join_smem_bytes(group_entries[g] + 1, 2, J.meta_cap, J.n_cand, join_warps);
            POMFRET_LAUNCH((join_kernel<true, uint16_t>), grid, join_threads, smem, st, JG);
        } else {
            JG.smem_tab_words = group_entries[g];
            const size_t smem = join_smem_bytes(group_entries[g] + 1, 4, J.meta_cap, J.n_cand, join_warps);
            POMFRET_LAUNCH((join_kernel<true, uint32_t>), grid, join_threads, smem, st, JG);
        }
        b->tm.launches++;
        if (n_side > 0) CK(cudaEventRecord(b->ev_side[(n_side - 1) % kSideStreams], st));
        n_side++;
    }
    for (int i = 0; i < std::min(n_side - 1, kSideStreams); i++) CK(cudaStreamWaitEvent(b->stream, b->ev_side[i], 0));
    CK(cudaEventRecord(b->ev[9], b->stream));
    CK(cudaGetLastError());
    b->stage = ST_JOINED;
    return POMFRET_GPU_OK;
}

// evaluate_separation1 (reference blockjoin.c:3894-3938) on the 2x2 table tabulated by the join kernel
static float evaluate_table(const int32_t t[4], int *join_dir) {
    const int b00 = t[0], b01 = t[1], b10 = t[2], b11 = t[3];
    const int buf[2][2] = {{b00, b01}, {b10, b11}};
    auto mn2 = [](int a, int c) { return a <= c ? a : c; };
    const int hard_cov_fail = mn2(b00, b01) > 15 || mn2(b10, b11) > 15;  // HARD_COV_THRESHOLD
    float scores[2] = {0, 0};
    int which_way = 0;
    for (int i = 0; i < 2; i++) {
        float mn, mx;
        if (buf[i][0] > buf[i][1]) { mn = (float)buf[i][1]; mx = (float)buf[i][0]; which_way = i == 0 ? which_way + 1 : which_way - 1; }
        else { mn = (float)buf[i][0]; mx = (float)buf[i][1]; which_way = i == 0 ? which_way - 1 : which_way + 1; }
        if (mn2(b00, b01) > 5 || mn2(b10, b11) > 5) { *join_dir = -9; return 1.0f; }  // HARD_CONTAMINATE_THRESHOLD
        if (mx == 0) { *join_dir = -9; return 1.0f; }
        mn = mn == 0 ? 1 : mn;
        if (mx / mn < 3) { *join_dir = -9; return 1.0f; }
        scores[i] = mx / mn;
    }
    double l, r, two;
    kt_fisher_exact(b00, b01, b10, b11, &l, &r, &two);
    if (two < 0.001 && !hard_cov_fail) { *join_dir = which_way; return scores[0] <= scores[1] ? scores[0] : scores[1]; }
    *join_dir = -9;
    return 1.0f;
}

int pomfret_gpu_batch_collect(pomfret_gpu_batch *b, pomfret_gpu_window_result *win, uint8_t *read_tags, int32_t *read_ids) {
    if (!b) return POMFRET_GPU_ERR_ARG;
    if (b->stage != ST_JOINED) return POMFRET_GPU_ERR_STATE;
    CK(cudaSetDevice(b->device));
    const size_t nr = b->h_reads.n, nw = b->h_win.n;
    int rc;
    if ((rc = b->h_state.resize(nw ? nw : 1))) return rc;
    b->host_tags_fwd.resize(nr + 1);
    b->host_rid.resize(nr + 1);
    if (nw) {
        CK(cudaMemcpyAsync(b->h_state.data(), b->d_state.p, nw * sizeof(WindowState), cudaMemcpyDeviceToHost, b->stream));
        CK(cudaMemcpyAsync(b->host_tags_fwd.data(), b->d_tags[0].p, nr, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaMemcpyAsync(b->host_rid.data(), b->d_r_id.p, nr * 4, cudaMemcpyDeviceToHost, b->stream));
        b->tm.bytes_d2h += nw * sizeof(WindowState) + nr * 5;
    }
    CK(cudaStreamSynchronize(b->stream));
    int first_err = 0;
    for (size_t w = 0; w < nw; w++) {
        const WindowState &S = b->h_state[w];
        const WindowRec &W = b->h_win[w];
        pomfret_gpu_window_result R;
        memset(&R, 0, sizeof(R));
        R.decision = R.join_fwd = R.join_bwd = -1;
        R.n_reads = (int32_t)S.n; R.n_reads_loaded = (int32_t)S.n_loaded;
        R.n_sites_fwd = R.n_sites_bwd = (int32_t)S.n_sites;
        R.n_left = (int32_t)S.n_left; R.n_left_strict = (int32_t)S.n_left_strict;
        R.n_right = (int32_t)S.n_right; R.n_right_strict = (int32_t)S.n_right_strict;
        R.status = S.status;
        R.score_fwd = R.score_bwd = 1.0f;
        if (S.status != 0 && !first_err) first_err = S.status;
        const bool ran = S.status == 0 && S.n > 0 && S.n_sites > 0;
        if (ran) {
            for (int t = 0; t < 4; t++) { R.table_fwd[t] = S.table[0][t]; R.table_bwd[t] = S.table[1][t]; }
            // haplotag_region2 with one permutation (blockjoin.c:4145-4156, 4188-4205)
            R.score_fwd = evaluate_table(S.table[0], &R.which_way_fwd);
            R.score_bwd = evaluate_table(S.table[1], &R.which_way_bwd);
            if (R.score_fwd >= 2 && R.which_way_fwd != 0) R.join_fwd = R.which_way_fwd > 0 ? 0 : 1;
            if (R.score_bwd >= 2 && R.which_way_bwd != 0) R.join_bwd = R.which_way_bwd > 0 ? 0 : 1;
            // blockjoin.c:4313-4320
            if (R.join_fwd != R.join_bwd || (R.join_fwd == -1 && R.join_bwd == -1)) R.decision = -1;
            else R.decision = R.join_fwd;
        }
        if (win) win[w] = R;
        for (uint32_t i = 0; i < W.n_reads; i++) {
            const size_t ri = (size_t)W.first_read + i;
            const int32_t id = b->host_rid[ri];
            if (read_ids) read_ids[ri] = id;
            if (read_tags) {
                uint8_t t = 255;
                if (id >= 0) {
                    if (!ran) t = (uint8_t)b->h_reads[ri].hp;                 // nothing touched the tags
                    else if (R.decision >= 0) t = b->host_tags_fwd[W.first_read + id];  // kept from the forward pass
                    else t = 2;                                               // set_all_as_unphased
                }
                read_tags[ri] = t;
            }
        }
    }
    // timings
    float ms;
    auto el = [&](int a, int c) { ms = 0; cudaEventElapsedTime(&ms, b->ev[a], b->ev[c]); return ms; };
    b->tm.h2d_ms = el(0, 1);
    b->tm.decode_ms = el(2, 3);
    if (nw) { b->tm.readset_ms = el(4, 5); b->tm.pileup_ms = el(5, 6); b->tm.methmer_ms = el(6, 7); b->tm.join_ms = el(8, 9); }
    b->tm.decode_bytes = b->alg_decode_bytes;
    uint64_t calls = 0, sites = 0;
    for (size_t w = 0; w < nw; w++) { calls += b->h_state[w].total_calls; sites += b->h_state[w].n_sites; }
    b->tm.decode_bytes += 5 * calls;
    b->tm.pileup_bytes = 5 * calls + 9 * sites;
    b->have_results = true;
    return first_err;
}

int pomfret_gpu_batch_timing(pomfret_gpu_batch *b, pomfret_gpu_timing *out) {
    if (!b || !out) return POMFRET_GPU_ERR_ARG;
    *out = b->tm;
    return POMFRET_GPU_OK;
}

// ---------------- debug / parity getters ----------------
static int dl(pomfret_gpu_batch *b, void *dst, const void *src, size_t bytes) {
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    if (bytes) CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int pomfret_gpu_debug_read_info(pomfret_gpu_batch *b, uint32_t read, uint32_t *status, uint32_t *n_calls, uint32_t *end_pos) {
    if (!b || read >= b->h_reads.n) return POMFRET_GPU_ERR_ARG;
    if (b->stage < ST_DECODED) return POMFRET_GPU_ERR_STATE;
    int rc;
    uint32_t v;
    if (status) { if ((rc = dl(b, &v, b->d_r_status.as<uint32_t>() + read, 4))) return rc; *status = v & 255u; }
    if (n_calls) { if ((rc = dl(b, &v, b->d_r_ncalls.as<uint32_t>() + read, 4))) return rc; *n_calls = v; }
    if (end_pos) { if ((rc = dl(b, &v, b->d_r_end.as<uint32_t>() + read, 4))) return rc; *end_pos = v; }
    return 0;
}

int pomfret_gpu_debug_get_calls(pomfret_gpu_batch *b, uint32_t read, uint32_t *pos, uint8_t *cat, uint32_t cap, uint32_t *n) {
    if (!b || read >= b->h_reads.n || !n) return POMFRET_GPU_ERR_ARG;
    if (b->stage < ST_DECODED) return POMFRET_GPU_ERR_STATE;
    uint32_t nc = 0, st = 0;
    int rc;
    if ((rc = dl(b, &nc, b->d_r_ncalls.as<uint32_t>() + read, 4))) return rc;
    if ((rc = dl(b, &st, b->d_r_status.as<uint32_t>() + read, 4))) return rc;
    if (!(st & RS_KEPT)) nc = 0;
    *n = nc;
    uint32_t m = std::min(nc, cap);
    const ReadRec &R = b->h_reads[read];
    if (pos && (rc = dl(b, pos, b->d_calls_pos.as<uint32_t>() + R.calls_off, (size_t)m * 4))) return rc;
    if (cat && (rc = dl(b, cat, b->d_calls_cat.as<uint8_t>() + R.calls_off, m))) return rc;
    return 0;
}

int pomfret_gpu_debug_get_sites(pomfret_gpu_batch *b, uint32_t window, int direction, uint32_t *real_pos, uint32_t *starts,
                                uint8_t *lens, uint32_t cap, uint32_t *n) {
    if (!b || window >= b->h_win.n || !n || direction < 0 || direction > 1) return POMFRET_GPU_ERR_ARG;
    if (b->stage < ST_PILED) return POMFRET_GPU_ERR_STATE;
    WindowState S;
    int rc;
    if ((rc = dl(b, &S, b->d_state.as<WindowState>() + window, sizeof(S)))) return rc;
    *n = S.n_sites;
    uint32_t m = std::min(S.n_sites, cap);
    const WindowRec &W = b->h_win[window];
    if (real_pos && (rc = dl(b, real_pos, b->d_site_pos.as<uint32_t>() + W.site_off, (size_t)m * 4))) return rc;
    if (starts && (rc = dl(b, starts, b->d_site_start[direction].as<uint32_t>() + W.site_off, (size_t)m * 4))) return rc;
    if (lens && (rc = dl(b, lens, b->d_site_len[direction].as<uint8_t>() + W.site_off, m))) return rc;
    return 0;
}

int pomfret_gpu_debug_get_mmrs(pomfret_gpu_batch *b, uint32_t read, int direction, uint32_t *mmr, uint32_t cap, uint32_t *n,
                               uint32_t *start_i) {
    if (!b || read >= b->h_reads.n || !n || direction < 0 || direction > 1) return POMFRET_GPU_ERR_ARG;
    if (b->stage < ST_PILED) return POMFRET_GPU_ERR_STATE;
    int rc;
    int32_t id = -1;
    if ((rc = dl(b, &id, b->d_r_id.as<int32_t>() + read, 4))) return rc;
    *n = 0;
    if (start_i) *start_i = 0;
    if (id < 0) return 0;
    const uint32_t w = b->h_read_win[read];
    if (w == 0xffffffffu) return 0;
    const uint32_t slot = b->h_win[w].first_read + (uint32_t)id;
    uint32_t cnt = 0, off = 0, st = 0;
    if ((rc = dl(b, &cnt, b->d_mm_n[direction].as<uint32_t>() + slot, 4))) return rc;
    if ((rc = dl(b, &off, b->d_mm_off[direction].as<uint32_t>() + slot, 4))) return rc;
    if ((rc = dl(b, &st, b->d_mm_start[direction].as<uint32_t>() + slot, 4))) return rc;
    *n = cnt;
    if (start_i) *start_i = st;
    uint32_t m = std::min(cnt, cap);
    if (mmr && m && (rc = dl(b, mmr, b->d_mmr_pool.as<uint32_t>() + off, (size_t)m * 4))) return rc;
    return 0;
}

int pomfret_gpu_debug_get_tags(pomfret_gpu_batch *b, int direction, uint8_t *tags) {
    if (!b || !tags || direction < 0 || direction > 1) return POMFRET_GPU_ERR_ARG;
    if (b->stage < ST_JOINED) return POMFRET_GPU_ERR_STATE;
    // slot-indexed (first_read + id) propagated tags of the direction
    return dl(b, tags, b->d_tags[direction].p, b->h_reads.n);
}

int pomfret_gpu_debug_get_tag_order(pomfret_gpu_batch *b, uint32_t window, int direction, uint32_t *ids, uint32_t cap, uint32_t *n) {
    if (!b || window >= b->h_win.n || !n || direction < 0 || direction > 1) return POMFRET_GPU_ERR_ARG;
    if (b->stage < ST_JOINED) return POMFRET_GPU_ERR_STATE;
    WindowState S;
    int rc;
    if ((rc = dl(b, &S, b->d_state.as<WindowState>() + window, sizeof(S)))) return rc;
    *n = S.n_order[direction];
    uint32_t m = std::min(S.n_order[direction], cap);
    if (ids && m && (rc = dl(b, ids, b->d_order[direction].as<uint32_t>() + b->h_win[window].first_read, (size_t)m * 4))) return rc;
    return 0;
}

}  // extern "C"

#include "engine_haptag.inc"
#include "engine_ingest.inc"
