// pomfret — drop-in front end for `pomfret methphase` / `pomfret report` (reference main.c:21-103) on top of
// the B200 engine (libpomfret_gpu.so).  Same options, inputs and output files.
#include <sys/resource.h>
#include <sys/time.h>
#include <cstdio>
#include <cstring>
#include "cli.h"
#include "methphase.h"

#define POMFRET_B200_VERSION "v0.1-r14-b200.1"

static double get_T() { struct timeval t; gettimeofday(&t, nullptr); return t.tv_sec + t.tv_usec / 1000000.0; }
static double get_U() { struct rusage s; getrusage(RUSAGE_SELF, &s); return (double)s.ru_maxrss / 1048576.0; }

int main(int argc, char **argv) {
    using namespace pomfret;
    fprintf(stderr, "[M::%s] pomfret %s\n", "main", POMFRET_B200_VERSION);
    fprintf(stderr, "[M::%s] CMD: ", "main");
    for (int i = 0; i < argc; i++) fprintf(stderr, "%s ", argv[i]);
    fprintf(stderr, "\n");
    const double T = get_T();
    int ret = 0;
    RunStats stats;
    if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help") || !strcmp(argv[1], "help")) {
        print_help_main();
        ret = 1;
    } else if (!strcmp(argv[1], "methphase") || !strcmp(argv[1], "report")) {
        Options opt;
        if (!parse_cli(argc - 1, argv + 1, &opt) || opt.is_help || !sancheck(&opt)) ret = 1;
        else if (!strcmp(argv[1], "methphase")) ret = run_methphase(opt, &stats);
        else if (opt.fn_vcf.empty()) { fprintf(stderr, "[E::%s] missing input: phasd vcf file.\n", "main"); ret = 1; }
        else run_report(opt, &stats);
    } else {
        fprintf(stderr, "[E::%s] unknown subcommand: %s\n", "main", argv[1]);
        print_help_main();
        ret = 1;
    }
    fprintf(stderr, "\n[M::%s] CMD: ", "main");
    for (int i = 0; i < argc; i++) fprintf(stderr, "%s ", argv[i]);
    fprintf(stderr, "\n");
    if (stats.n_windows || stats.n_haptag_reads)
        fprintf(stderr, "[M::%s] windows %llu, reads %llu, bases %llu; haplotagged reads %llu; load %.2fs, gpu %.2fs, haptag %.2fs\n", "main",
                (unsigned long long)stats.n_windows, (unsigned long long)stats.n_reads, (unsigned long long)stats.n_bases,
                (unsigned long long)stats.n_haptag_reads, stats.t_load, stats.t_gpu, stats.t_haptag);
    if (stats.n_shared || stats.n_ingest_bytes)
        fprintf(stderr, "[M::%s] %llu window slots reuse a record decoded for another window; %.1f MB of BGZF blocks inflated on the device\n", "main",
                (unsigned long long)stats.n_shared, stats.n_ingest_bytes / 1e6);
    fprintf(stderr, "[M::%s] used: %.1fs, peak RSS %.1fGiB\n", "main", get_T() - T, get_U());
    return ret;
}
