// pomfret-synth: command-line front end of the synthetic data generator.
//   pomfret-synth -o prefix [-s seed] [-c coverage] [-C name:len[:beg-end]]... [--untagged] [--preset NAME]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "synth.h"

static void usage() {
    fprintf(stderr,
            "Usage: pomfret-synth -o prefix [options]\n"
            "  -s INT        seed [20]\n"
            "  -c FLOAT      coverage [30]\n"
            "  -C SPEC       contig as name:length[:beg-end]; repeatable [chr20:64444167]\n"
            "  -l INT        deflate level 0-9 [1]\n"
            "  -F INT        filler header contigs before the first real one [0]\n"
            "  --untagged    do not write HP/PS tags (for methphase -u)\n"
            "  --qual        write pseudo-random base qualities instead of 0xff\n"
            "  --implicit F  fraction of non-CpG C positions that also get a C+m call [0]\n"
            "  --listed F    fraction of CpG cytosines present in the MM list [1]\n"
            "  --err F       per-base error rate [0.01]\n"
            "  --de-cap F    upper bound of the reported de tag [none]\n"
            "  --frac-meth F / --frac-unmeth F  CpG sites methylated / unmethylated on both haplotypes [0.70 / 0.15];\n"
            "                the rest is haplotype specific\n"
            "  --ml-flip F   probability that a call shows the wrong state [0.08]\n"
            "  --hp-drop F   fraction of reads inside a phase block that get no HP tag [0.1]\n"
            "  --vcf-in FILE take variants, phase sets and the header's contig list from FILE; -C names a contig of it\n"
            "  --readlen F   mean read length [20000]\n"
            "  --gap A-B     phase-block gap length range [20000-150000]\n"
            "  --block F     median phase-block length [500000]\n");
}

int main(int argc, char **argv) {
    synth_config cfg;
    pomfret_synth_default_config(&cfg);
    std::string prefix;
    std::vector<std::string> names;
    std::vector<int64_t> lens, begs, ends;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char *what) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "missing argument for %s\n", what); exit(1); }
            return argv[++i];
        };
        if (a == "-o") prefix = need("-o");
        else if (a == "-s") cfg.seed = strtoull(need("-s"), nullptr, 10);
        else if (a == "-c") cfg.coverage = atof(need("-c"));
        else if (a == "-l") cfg.compress_level = atoi(need("-l"));
        else if (a == "-F") cfg.n_header_contigs_before = atoi(need("-F"));
        else if (a == "--untagged") cfg.tagged = 0;
        else if (a == "--qual") cfg.qual_mode = 1;
        else if (a == "--implicit") cfg.frac_noncpg_calls = atof(need("--implicit"));
        else if (a == "--listed") cfg.frac_cpg_listed = atof(need("--listed"));
        else if (a == "--err") cfg.err_rate = atof(need("--err"));
        else if (a == "--de-cap") cfg.de_cap = atof(need("--de-cap"));
        else if (a == "--frac-meth") cfg.frac_meth = atof(need("--frac-meth"));
        else if (a == "--frac-unmeth") cfg.frac_unmeth = atof(need("--frac-unmeth"));
        else if (a == "--ml-flip") cfg.ml_flip = atof(need("--ml-flip"));
        else if (a == "--vcf-in") cfg.vcf_in = need("--vcf-in");
        else if (a == "--hp-drop") cfg.hp_drop = atof(need("--hp-drop"));
        else if (a == "--readlen") cfg.read_len_mean = atof(need("--readlen"));
        else if (a == "--block") cfg.block_len_median = atof(need("--block"));
        else if (a == "--gap") {
            long long x, y;
            if (sscanf(need("--gap"), "%lld-%lld", &x, &y) != 2) { usage(); return 1; }
            cfg.gap_min = x; cfg.gap_max = y;
        } else if (a == "-C") {
            std::string spec = need("-C");
            size_t c1 = spec.find(':');
            if (c1 == std::string::npos) { usage(); return 1; }
            names.push_back(spec.substr(0, c1));
            size_t c2 = spec.find(':', c1 + 1);
            lens.push_back(atoll(spec.substr(c1 + 1, c2 == std::string::npos ? std::string::npos : c2 - c1 - 1).c_str()));
            long long x = 0, y = 0;
            if (c2 != std::string::npos) sscanf(spec.c_str() + c2 + 1, "%lld-%lld", &x, &y);
            begs.push_back(x);
            ends.push_back(y);
        } else { usage(); return 1; }
    }
    if (prefix.empty()) { usage(); return 1; }
    if (names.empty()) { names.push_back("chr20"); lens.push_back(64444167); begs.push_back(0); ends.push_back(0); }
    std::vector<const char *> np;
    for (auto &s : names) np.push_back(s.c_str());
    int rc = pomfret_synth_write(&cfg, np.data(), lens.data(), begs.data(), ends.data(), (int)np.size(), prefix.c_str());
    if (rc != 0) fprintf(stderr, "pomfret-synth failed: %d\n", rc);
    return rc ? 1 : 0;
}
