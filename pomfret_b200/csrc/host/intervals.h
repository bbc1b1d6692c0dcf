// Phase-block bookkeeping of the host front end (hot-path scope rows a16/a17 of SURVEY.md §8):
// interval loading from VCF / GTF / TSV, gap merging, decision lifting, flips and new phase blocks.
// Behaviour follows the reference (blockjoin.c:1300-1430, 1977-2218, 2250-2361) including the quirks
// listed in SURVEY.md App. A.8, because the output files must be bit-exact.
#ifndef POMFRET_HOST_INTERVALS_H
#define POMFRET_HOST_INTERVALS_H
#include <cstdint>
#include <functional>
#include <string>
#include <vector>
#include "pomfret_gpu.h"

namespace pomfret {

enum class IntervalFormat { GTF = 0, VCF = 1, TSV = 2 };

struct U32Pair { uint32_t s, e; };

// ranges_t (blockjoin.c:1178-1189).  `starts`/`ends` keep their pre-merge length and stale tail on
// purpose: lift_decisions() iterates over the pre-merge count (blockjoin.c:2215, 2257).
struct Ranges {
    uint32_t abs_start = 0, abs_end = 0;
    std::vector<U32Pair> dropped;
    std::vector<uint32_t> starts, ends;
    size_t n = 0;            // starts.n / ends.n
    std::vector<int> decisions;
    size_t n_decisions = 0;  // decisions.n
    std::vector<U32Pair> raw;
    std::vector<int> decisions_onraw, flips_onraw;
    std::vector<U32Pair> phaseblocks;
};

// Known phased variants of one contig (vvar_t filled by insert_variant_from_vcf_line, blockjoin.c:1432-1543)
struct KnownVariants {
    std::vector<pomfret_gpu_variant> vars;
    std::vector<uint8_t> bases;
    void clear() { vars.clear(); bases.clear(); }
};

struct Storage {
    std::vector<std::string> ref_names;
    std::vector<Ranges> ranges;
};

// Called whenever the loader leaves a contig (and at EOF) with the variants collected for it, in the
// order the reference triggers pre_haplotagging_read_in_one_ref (blockjoin.c:2065-2081, 2149-2155).
using ContigVariantsFn = std::function<void(const std::string &chrom, KnownVariants &vars, bool at_eof)>;

// load_intervals_from_file (blockjoin.c:1977-2176).  on_contig may be empty.
// Returns false if the file cannot be opened; *fatal receives a message for conditions on which the
// reference exits (unsorted VCF).
bool load_intervals(const std::string &fn, IntervalFormat fmt, Storage *st, const ContigVariantsFn &on_contig,
                    std::string *fatal);

// insert_variant_from_vcf_line on an already tokenised record (exposed for tests)
bool variant_from_vcf_fields(const char *ref, const char *alt, const char *format, const char *sample, uint32_t pos1,
                             KnownVariants *out);

void store_raw_intervals(Ranges *r);                  // blockjoin.c:2178-2188
void merge_close_intervals(Ranges *r, int threshold); // blockjoin.c:2190-2218
void lift_decisions(Storage *st);                     // blockjoin.c:2250-2310
void make_flips_onraw(Storage *st);                   // blockjoin.c:2312-2324
void generate_new_phase_blocks(Storage *st);          // blockjoin.c:2326-2361 with use_raw = 1

}  // namespace pomfret
#endif
