// Command line of `pomfret methphase` / `pomfret report`: same options, defaults and sanity checks as the
// reference (cli.c:28-74, 120-243, 245-325).
#ifndef POMFRET_HOST_CLI_H
#define POMFRET_HOST_CLI_H
#include <string>

namespace pomfret {

struct Options {
    bool is_help = false;
    int threads = 1, threads_bam = 1;
    int lo = 100, hi = 156;
    std::string fn_tsv, fn_gtf, fn_vcf, fn_bam;
    bool bam_needs_haplotagging = false, write_bam_input_haplotagging = false;
    std::string output_prefix = "pomfret";
    int readlen_threshold = 15000, mapq = 10, k = 3, k_span = 5000;
    int cov = 0;                 // the reference leaves this uninitialised without -c
    int cov_for_selection = -1, n_candidates_per_iter = 15;
    bool do_output_bam = false, do_output_tsv = false, write_debug_files = false;
    int chunk_size = 50000, chunk_stride = 1000000;
    int verbose = 0;
    // additions of this implementation (do not change results)
    int gpus = 0;                // 0 = not given: one device per 4 GiB of BAM; "all" = every visible device
    int windows_per_batch = 8;
};

void print_help_main();
void print_help_methphase(const Options &o);
// parse_cli: returns false on a parse error (message already printed)
bool parse_cli(int argc, char **argv, Options *o);
// sancheck_cliopt: returns false if the run must stop
bool sancheck(Options *o);

}  // namespace pomfret
#endif
