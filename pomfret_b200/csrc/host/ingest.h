// Host half of the compressed ingest (include/pomfret_gpu.h, "(a') compressed ingest"): which BGZF blocks a region
// query needs, read as they lie in the file — no inflate, no record parsing on the host.  The index chunks are the
// ones htslib's iterator would walk (sam_itr_queryi: bins of the region, linear-index cut, merged), so the device
// sees exactly the records sam_itr_next would return for load_reads_given_interval (blockjoin.c:1061-1070).
#ifndef POMFRET_HOST_INGEST_H
#define POMFRET_HOST_INGEST_H
#include <cstdint>
#include <string>
#include <vector>
#include "htslib/sam.h"
#include "pomfret_gpu.h"

namespace pomfret {

struct IngestPlan {
    struct Range { uint64_t file_off, bytes, comp_off; uint64_t vbeg, vend; uint32_t run; int32_t tid; uint32_t end0; };
    std::vector<Range> ranges;   // one per index chunk
    size_t comp_bytes = 0;       // upper bound of the compressed bytes (every range is padded by one block)
    std::vector<pomfret_gpu_bgzf_block> blocks;
    std::vector<pomfret_gpu_bgzf_stream> streams;
    std::vector<uint32_t> stream_run;  // the run (region query) a stream belongs to
    void clear() { ranges.clear(); comp_bytes = 0; blocks.clear(); streams.clear(); stream_run.clear(); }
};

// add the chunks of the query [beg0, end0) on target tid; returns false if the query cannot be made
bool ingest_plan_region(hts_idx_t *idx, int tid, int64_t beg0, int64_t end0, uint32_t run, uint64_t file_size, IngestPlan *plan);
// read the ranges into `comp` (comp_bytes large), then walk the block headers there: blocks[], streams[]
bool ingest_read(int fd, IngestPlan *plan, uint8_t *comp, std::string *err);

}  // namespace pomfret
#endif
