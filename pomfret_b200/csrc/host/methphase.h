// Drivers of the host front end: `pomfret methphase` (main_blockjoin, reference blockjoin.c:4643-4735 with
// blockjoin_parallel 4428-4603) and `pomfret report` (main_methreport, 4908-5097).  I/O, option handling,
// interval bookkeeping and writers run on the host; every window and every -u read goes through the
// pomfret_gpu_* C ABI.
#ifndef POMFRET_HOST_METHPHASE_H
#define POMFRET_HOST_METHPHASE_H
#include <cstdint>
#include <string>
#include <vector>
#include "cli.h"
#include "writers.h"

namespace pomfret {

struct RunStats {
    uint64_t n_windows = 0, n_reads = 0, n_bases = 0;      // window slots / bases of the distinct records behind them
    uint64_t n_ingest_bytes = 0;                           // compressed bytes shipped to the device (compressed ingest)
    uint64_t n_shared = 0;                                 // slots that reuse a record staged and decoded for an earlier window
    uint64_t n_haptag_reads = 0, n_haptag_bases = 0;       // records fed to the -u haplotagger
    double t_load = 0, t_gpu = 0, t_haptag = 0, t_total = 0;
    double t_phase[12] = {};                               // per-phase wall times of the compressed-ingest loader (diagnostics)
};

int run_methphase(const Options &opt, RunStats *stats);
int run_report(const Options &opt, RunStats *stats);

// estimate_read_coverage_dirtyfast (blockjoin.c:951-1040); one entry per BAM header target
std::vector<int> estimate_read_coverage(const std::string &fn_bam);

}  // namespace pomfret
#endif
