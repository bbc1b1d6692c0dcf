#include "cli.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace pomfret {

void print_help_main() {
    fprintf(stderr, "Usage: pomfret <subcommand> [options]\n");
    fprintf(stderr, "Subcommands:\n");
    fprintf(stderr, "  methphase  Given aligned reads with methylation calls in bam\n");
    fprintf(stderr, "             and exiting phase blocks, try to use methylation to\n");
    fprintf(stderr, "             phase the unphased regions.\n");
    fprintf(stderr, "  report     Given aligned reads in bam and a phased vcf, sample\n");
    fprintf(stderr, "             intervals within phase blocks, pretend they are phase gaps\n");
    fprintf(stderr, "             and report whether meth-phasing would generate correct \n");
    fprintf(stderr, "             phase block joining decisions.\n");
}

void print_help_methphase(const Options &o) {
    fprintf(stderr, "Usage: pomfret methphase -o out_prefix --vcf phased.vcf[.gz] [...] reads.bam 2>log\n");
    fprintf(stderr, "Options:\n");
    fprintf(stderr, "  bam    [pos] Aligned reads. Must be sorted and has index. If reads are not\n");
    fprintf(stderr, "               haplotagged, supply -u and provide vcf (via --vcf).\n");
    fprintf(stderr, "  -h,--help [   ] Display this message.\n");
    fprintf(stderr, "  -c     [opt] Read coverage (total, not per-haplotap). Will infer if not supplied.\n");
    fprintf(stderr, "  -o     [opt] Name prefix of output files. [%s]\n", o.output_prefix.c_str());
    fprintf(stderr, "  --vcf  [opt] Input, sorted vcf file containing phased variants.\n");
    fprintf(stderr, "               Either vcf, gtf or tsv need to be present. Plain or gz'd.\n");
    fprintf(stderr, "  --gtf  [opt] Input, sorted gtf file of prescribed phase blocks.\n"
                    "               If present, overrides phase blocks defined by vcf.\n"
                    "               Plain or gz'd.\n");
    fprintf(stderr, "  --tsv  [opt] Input, sorted 3-column tsv file of prescribed phase blocks:\n"
                    "               reference name, start, end. Plain or gz'd.\n"
                    "               If present, overrides both gtf and vcf.\n");
    fprintf(stderr, "  -u,--bam-is-untagged [opt] If present, will haplotag reads \n"
                    "               with phased variants in vcf first. --vcf must be \n"
                    "               supplied. Ignores any haptags present in the bam.\n"
                    "               All variants supplies by the vcf will be used as evidences.\n");
    fprintf(stderr, "  -t     [opt] Number of threads to use. [%d]\n", o.threads);
    fprintf(stderr, "  --gpus [opt] Number of B200 devices to shard contigs over. [of the visible ones, one per 4 GiB of BAM]\n");
    fprintf(stderr, "Note: Inputs may need to be opened or read for more than once.\n");
}

namespace {
struct LongOpt { const char *name; bool has_arg; int val; };
const LongOpt kLong[] = {
    {"lo", true, 301}, {"hi", true, 302}, {"gtf", true, 304}, {"vcf", true, 306}, {"mapq", true, 307}, {"tsv", true, 308},
    {"write-bam", false, 309}, {"output-tsv", false, 310}, {"bam-threads", true, 311}, {"bam-is-untagged", false, 312},
    {"write-input-tagging", false, 313}, {"chunk-size", true, 314}, {"chunk-stride", true, 315}, {"help", false, 400},
    {"dbg", false, 401}, {"gpus", true, 501}, {"windows-per-batch", true, 502}, {nullptr, false, 0}};
const char *kShort = "vhuUo:k:L:l:c:n:t:T:";
}  // namespace

bool parse_cli(int argc, char **argv, Options *o) {
    if (argc == 1) { print_help_methphase(*o); return false; }
    std::vector<const char *> positional;
    auto apply = [&](int c, const char *arg) {
        switch (c) {
        case 'v': o->verbose++; break;
        case 'h': case 400: print_help_methphase(*o); o->is_help = true; break;
        case 't': o->threads = atoi(arg); o->threads_bam = o->threads; break;
        case 'o': o->output_prefix = arg; break;
        case 'k': o->k = atoi(arg); break;
        case 'l': o->k_span = atoi(arg); break;
        case 'L': o->readlen_threshold = atoi(arg); break;
        case 'c': {
            int cov = atoi(arg);
            o->cov = cov; o->cov_for_selection = cov / 10; o->n_candidates_per_iter = cov / 4;
            break;
        }
        case 'n': o->n_candidates_per_iter = atoi(arg); break;
        case 301: o->lo = atoi(arg); break;
        case 302: o->hi = atoi(arg); break;
        case 304: o->fn_gtf = arg; break;
        case 306: o->fn_vcf = arg; break;
        case 307: o->mapq = atoi(arg); break;
        case 308: o->fn_tsv = arg; break;
        case 309: o->do_output_bam = true; break;
        case 310: o->do_output_tsv = true; break;
        case 401: o->write_debug_files = true; break;
        case 'T': case 311: o->threads_bam = atoi(arg); break;
        case 312: case 'u': o->bam_needs_haplotagging = true; break;
        case 313: case 'U': o->write_bam_input_haplotagging = true; break;
        case 314: o->chunk_size = atoi(arg); break;
        case 315: o->chunk_stride = atoi(arg); break;
        case 501: o->gpus = strcmp(arg, "all") == 0 ? 1024 : atoi(arg); break;  // (absent: chosen from the size of the input)
        case 502: o->windows_per_batch = atoi(arg); break;
        default: break;
        }
    };
    bool only_positional = false;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (only_positional || a[0] != '-' || a[1] == 0) { positional.push_back(a); continue; }
        if (a[1] == '-') {
            if (a[2] == 0) { only_positional = true; continue; }
            const char *eq = strchr(a + 2, '=');
            size_t nl = eq ? (size_t)(eq - (a + 2)) : strlen(a + 2);
            const LongOpt *hit = nullptr;
            int n_hit = 0;
            for (const LongOpt *l = kLong; l->name; l++) {
                if (strlen(l->name) == nl && strncmp(l->name, a + 2, nl) == 0) { hit = l; n_hit = 1; break; }  // exact
                if (strncmp(l->name, a + 2, nl) == 0) { hit = l; n_hit++; }                                    // prefix
            }
            if (!hit || n_hit != 1) { fprintf(stderr, "[E::%s] unknown option argument in \"%s\"\n", "parse_cli", a); return false; }
            const char *arg = nullptr;
            if (hit->has_arg) {
                if (eq) arg = eq + 1;
                else if (i + 1 < argc) arg = argv[++i];
                else { fprintf(stderr, "[E::%s] missing option argument in \"%s\"\n", "parse_cli", a); return false; }
            }
            apply(hit->val, arg);
            continue;
        }
        for (const char *p = a + 1; *p; p++) {
            const char *s = strchr(kShort, *p);
            if (!s || *p == ':') { fprintf(stderr, "[E::%s] unknown option argument in \"%s\"\n", "parse_cli", a); return false; }
            if (s[1] == ':') {
                const char *arg = p[1] ? p + 1 : (i + 1 < argc ? argv[++i] : nullptr);
                if (!arg) { fprintf(stderr, "[E::%s] missing option argument in \"%s\"\n", "parse_cli", a); return false; }
                apply(*p, arg);
                break;
            }
            apply(*p, nullptr);
        }
    }
    if (positional.size() > 1) {
        fprintf(stderr, "[TODO::%s] too many positional arguments; multi bam input not impl'd yet\n", "parse_cli");
        fprintf(stderr, "[E::%s] multiple bam input is not supported.\n", "parse_cli");
        exit(1);
    }
    if (!positional.empty()) o->fn_bam = positional[0];
    putchar('\n');
    return true;
}

bool sancheck(Options *o) {
    if (o->threads <= 0) { fprintf(stderr, "[W::%s] invalid thread number (%d), clipped to 1\n", "sancheck_cliopt", o->threads); o->threads = 1; }
    if (o->threads_bam <= 0) { fprintf(stderr, "[W::%s] invalid bam thread number(%d), clipped to 1\n", "sancheck_cliopt", o->threads_bam); o->threads_bam = 1; }
    if (o->lo < 0) { fprintf(stderr, "[E::%s] lower threshold for mod call quality is too low (%d)\n", "sancheck_cliopt", o->lo); return false; }
    if (o->lo > 127) { fprintf(stderr, "[E::%s] lower threshold for mod call quality is too high (%d)\n", "sancheck_cliopt", o->lo); return false; }
    if (o->hi > 255) { fprintf(stderr, "[E::%s] upper threshold for mod call quality is too high (%d)\n", "sancheck_cliopt", o->hi); return false; }
    if (o->hi <= 127) { fprintf(stderr, "[E::%s] upper threshold for mod call quality is too low (%d)\n", "sancheck_cliopt", o->hi); return false; }
    if (o->readlen_threshold < 0) o->readlen_threshold = 0;
    if (o->mapq > 60) fprintf(stderr, "[W::%s] mapq seems too high, proceed anyways\n", "sancheck_cliopt");
    if (o->mapq < 0) o->mapq = 0;
    if (o->k <= 0) { fprintf(stderr, "[W::%s] clipping mether k to 1\n", "sancheck_cliopt"); o->k = 1; }
    if (o->k_span <= 0) { fprintf(stderr, "[W::%s] clipping mether span to 1\n", "sancheck_cliopt"); o->k_span = 1; }
    if (o->cov_for_selection <= 0) fprintf(stderr, "[M::%s] read coverage not provided, will estimate.\n", "sancheck_cliopt");
    if (o->n_candidates_per_iter <= 0) { fprintf(stderr, "[W::%s] clipping candidate per iter to 1\n", "sancheck_cliopt"); o->n_candidates_per_iter = 1; }
    if (o->n_candidates_per_iter < 5) fprintf(stderr, "[W::%s] number of candidates per iter might be too low\n", "sancheck_cliopt");
    if (o->fn_gtf.empty() && o->fn_tsv.empty() && o->fn_vcf.empty()) { fprintf(stderr, "[E::%s] gtf, tsv and vcf cannot all be absent\n", "sancheck_cliopt"); return false; }
    if ((!o->fn_gtf.empty()) + (!o->fn_tsv.empty()) + (!o->fn_vcf.empty()) > 1)
        fprintf(stderr, "[M::%s] multiple phase block files. Will not resolve conflict, only total override (order is always: tsv > gtf > vcf)\n", "sancheck_cliopt");
    if (o->bam_needs_haplotagging && o->fn_vcf.empty()) { fprintf(stderr, "[E::%s] input bam was flagged unhaplotagged, but vcf is missing.\n", "sancheck_cliopt"); return false; }
    if (o->fn_bam.empty()) { fprintf(stderr, "[E::%s] missing bam file\n", "sancheck_cliopt"); return false; }
    if (o->output_prefix.empty()) { fprintf(stderr, "[E::%s] no output prefix given\n", "sancheck_cliopt"); return false; }
    if (o->output_prefix.back() == '/') {
        fprintf(stderr, "[W::%s] output prefix has trailing '/', stripping them.\n", "sancheck_cliopt");
        while (!o->output_prefix.empty() && o->output_prefix.back() == '/') o->output_prefix.pop_back();
    }
    if (o->output_prefix.empty()) { fprintf(stderr, "[E::%s] no output prefix given\n", "sancheck_cliopt"); return false; }
    if (o->chunk_size <= 0) { fprintf(stderr, "[E::%s] invalid chunk size\n", "sancheck_cliopt"); return false; }
    if (o->chunk_stride <= 0) { fprintf(stderr, "[E::%s] invalid chunk stride\n", "sancheck_cliopt"); return false; }
    return true;
}

}  // namespace pomfret
