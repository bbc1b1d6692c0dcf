#include "gpu_api.h"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace pomfret {

GpuApi &gpu_api() {
    static GpuApi api;
    return api;
}

static std::string dir_of_this_library() {
    Dl_info info;
    if (dladdr((void *)&gpu_api, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.rfind('/');
        if (s != std::string::npos) return p.substr(0, s);
    }
    return ".";
}

bool GpuApi::load(std::string *err) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (handle) return true;
    std::string tried;
    const char *env = getenv("POMFRET_GPU_LIB");
    std::string here = dir_of_this_library();
    std::string cands[4] = {env ? env : "", here + "/libpomfret_gpu.so", here + "/../lib/libpomfret_gpu.so", "libpomfret_gpu.so"};
    for (const std::string &c : cands) {
        if (c.empty()) continue;
        handle = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (handle) break;
        tried += "\n  " + c + ": " + dlerror();
    }
    if (!handle) {
        if (err) *err = "cannot load libpomfret_gpu.so (the CUDA engine; there is no CPU fallback):" + tried;
        return false;
    }
#define BIND(field, sym)                                                        \
    *(void **)(&field) = dlsym(handle, sym);                                    \
    if (!field) { if (err) *err = std::string("missing symbol ") + sym; dlclose(handle); handle = nullptr; return false; }
    BIND(init, "pomfret_gpu_init") BIND(destroy, "pomfret_gpu_destroy") BIND(strerror, "pomfret_gpu_strerror")
    BIND(device_count, "pomfret_gpu_device_count") BIND(batch_begin, "pomfret_gpu_batch_begin")
    BIND(batch_reset, "pomfret_gpu_batch_reset") BIND(batch_add_read, "pomfret_gpu_batch_add_read")
    BIND(batch_add_reads, "pomfret_gpu_batch_add_reads") BIND(batch_add_reads_shared, "pomfret_gpu_batch_add_reads_shared")
    BIND(batch_add_windows, "pomfret_gpu_batch_add_windows") BIND(host_register, "pomfret_gpu_host_register")
    BIND(host_unregister, "pomfret_gpu_host_unregister") BIND(batch_add_reads_device, "pomfret_gpu_batch_add_reads_device")
    BIND(batch_ingest_buffer, "pomfret_gpu_batch_ingest_buffer") BIND(batch_ingest_bgzf, "pomfret_gpu_batch_ingest_bgzf")
    BIND(batch_ingest_records, "pomfret_gpu_batch_ingest_records") BIND(batch_ingest_qname, "pomfret_gpu_batch_ingest_qname")
    BIND(batch_ingest_coverage, "pomfret_gpu_batch_ingest_coverage") BIND(batch_ingest_retag, "pomfret_gpu_batch_ingest_retag")
    BIND(variant_votes, "pomfret_gpu_variant_votes") BIND(batch_add_window, "pomfret_gpu_batch_add_window") BIND(batch_submit, "pomfret_gpu_batch_submit")
    BIND(decode, "pomfret_gpu_decode") BIND(haptag, "pomfret_gpu_haptag") BIND(pileup, "pomfret_gpu_pileup")
    BIND(join, "pomfret_gpu_join") BIND(batch_collect, "pomfret_gpu_batch_collect")
    BIND(batch_collect_haptags, "pomfret_gpu_batch_collect_haptags") BIND(batch_end, "pomfret_gpu_batch_end")
    BIND(batch_timing, "pomfret_gpu_batch_timing")
#undef BIND
    return true;
}

}  // namespace pomfret
