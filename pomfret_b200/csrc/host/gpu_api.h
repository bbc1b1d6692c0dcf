// Run-time binding of the pomfret_gpu_* C ABI (include/pomfret_gpu.h).  The host front end has no
// compute path of its own: if libpomfret_gpu.so cannot be loaded or finds no device, the run fails.
#ifndef POMFRET_HOST_GPU_API_H
#define POMFRET_HOST_GPU_API_H
#include <string>
#include "pomfret_gpu.h"

namespace pomfret {

struct GpuApi {
    void *handle = nullptr;
    int (*init)(pomfret_gpu_ctx **, const int *, int, int) = nullptr;
    void (*destroy)(pomfret_gpu_ctx *) = nullptr;
    const char *(*strerror)(int) = nullptr;
    int (*device_count)(void) = nullptr;
    int (*batch_begin)(pomfret_gpu_ctx *, int, int, pomfret_gpu_batch **) = nullptr;
    int (*batch_reset)(pomfret_gpu_batch *) = nullptr;
    int (*batch_add_read)(pomfret_gpu_batch *, const pomfret_gpu_read_desc *) = nullptr;
    int (*batch_add_reads)(pomfret_gpu_batch *, const pomfret_gpu_read_desc *, uint32_t) = nullptr;
    int (*batch_add_reads_shared)(pomfret_gpu_batch *, const pomfret_gpu_read_desc *, uint32_t, const int64_t *) = nullptr;
    int (*batch_add_window)(pomfret_gpu_batch *, uint32_t, uint32_t, uint32_t, uint32_t) = nullptr;
    int (*batch_add_windows)(pomfret_gpu_batch *, const uint32_t *, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t) = nullptr;
    int (*batch_add_reads_device)(pomfret_gpu_batch *, const pomfret_gpu_read_desc *, uint32_t, const int64_t *) = nullptr;
    int (*batch_ingest_buffer)(pomfret_gpu_batch *, size_t, void **) = nullptr;
    int (*batch_ingest_bgzf)(pomfret_gpu_batch *, const void *, size_t, const pomfret_gpu_bgzf_block *, uint32_t, const pomfret_gpu_bgzf_stream *,
                             uint32_t, const pomfret_gpu_ingest_filter *, uint32_t *) = nullptr;
    int (*batch_ingest_records)(pomfret_gpu_batch *, pomfret_gpu_sliced_record *, uint32_t) = nullptr;
    int (*batch_ingest_qname)(pomfret_gpu_batch *, uint32_t, char *, uint32_t) = nullptr;
    int (*batch_ingest_coverage)(pomfret_gpu_batch *, uint32_t, uint32_t, uint32_t, uint64_t *) = nullptr;
    int (*batch_ingest_retag)(pomfret_gpu_batch *, const uint64_t *, const uint8_t *, uint64_t, void *) = nullptr;
    int (*variant_votes)(pomfret_gpu_batch *, const uint32_t *, uint32_t, const uint8_t *, int32_t *) = nullptr;
    int (*host_register)(pomfret_gpu_ctx *, void *, size_t) = nullptr;
    int (*host_unregister)(pomfret_gpu_ctx *, void *) = nullptr;
    int (*batch_submit)(pomfret_gpu_batch *) = nullptr;
    int (*decode)(pomfret_gpu_batch *, uint8_t, uint8_t) = nullptr;
    int (*haptag)(pomfret_gpu_batch *, const pomfret_gpu_variant *, uint32_t, const uint8_t *, uint32_t, const uint32_t *) = nullptr;
    int (*pileup)(pomfret_gpu_batch *, const pomfret_gpu_config *) = nullptr;
    int (*join)(pomfret_gpu_batch *, const pomfret_gpu_config *) = nullptr;
    int (*batch_collect)(pomfret_gpu_batch *, pomfret_gpu_window_result *, uint8_t *, int32_t *) = nullptr;
    int (*batch_collect_haptags)(pomfret_gpu_batch *, uint8_t *, int32_t *) = nullptr;
    void (*batch_end)(pomfret_gpu_batch *) = nullptr;
    int (*batch_timing)(pomfret_gpu_batch *, pomfret_gpu_timing *) = nullptr;
    // Loads $POMFRET_GPU_LIB or libpomfret_gpu.so next to this library / executable.
    bool load(std::string *err);
};

GpuApi &gpu_api();

}  // namespace pomfret
#endif
