// Output surface of `pomfret methphase` (hot-path scope row a17): <prefix>.mp.gtf / .tsv / .vcf / .bam.
// Restates the reference writers (blockjoin.c:2365-2473, 2475-2692, 2695-2755, 2758-2988, 2990-3103) so that
// the files are byte-identical; quirks are kept and marked.
#ifndef POMFRET_HOST_WRITERS_H
#define POMFRET_HOST_WRITERS_H
#include <cstdint>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>
#include "intervals.h"
#include "loader.h"

namespace pomfret {

using TagMap = std::unordered_map<std::string, int>;

struct PhaseState {
    Storage st;
    TagMap qname2haptag;      // tags from meth phasing (first insert wins, blockjoin.c:4408-4423, 4579-4595)
    TagMap qname2haptag_raw;  // -u: tags from the VCF haplotagger
    bool stores_raw_tag = false;
    std::vector<std::unordered_map<uint32_t, uint32_t>> varphase_in_dropped;  // per contig: 0-based pos -> hap of REF
};

void output_gtf(const PhaseState &ps, const std::string &prefix);  // blockjoin.c:2721-2755
void output_tsv(const PhaseState &ps, const std::string &prefix);  // blockjoin.c:2695-2719
void output_debug_read2tag(const PhaseState &ps, const std::string &prefix);  // blockjoin.c:2223-2248 (hash order differs)

// recover_variant_phase_in_dropped_intervals (blockjoin.c:2618-2692)
// `on_device` (may be empty): evaluates one interval on the device instead of the host loop below it
// (recover_variant_phase_in_one_interval, blockjoin.c:2475-2600)
using RecoverIntervalFn = std::function<void(const std::string &refname, uint32_t start, uint32_t end, const std::vector<uint32_t> &poss,
                                             std::unordered_map<uint32_t, uint32_t> *pos2hap)>;
int recover_variant_phase_in_dropped_intervals(PhaseState *ps, const std::string &fn_bam, const std::string &fn_vcf,
                                               const RecoverIntervalFn &on_device = RecoverIntervalFn());
// output_modify_vcf (blockjoin.c:2909-2988); returns 0 or 1 on a fatal header problem
int output_modify_vcf(const std::string &fn_vcf, const PhaseState &ps, const std::string &prefix);
// output_modify_bam + sam_index_build3 (blockjoin.c:3022-3103, 4714-4731)
int output_modify_bam(const std::string &fn_bam, const PhaseState &ps, const std::string &fn_out);
// the tag a record gets there (blockjoin.c:3056-3092), stateful over the records in file order; hp_of_record is
// get_hp_from_aln of the record (used unless the run stores raw tags)
struct RetagCursor { int prev_unphased_idx = 1, prev_tid = 0, need_flip = 0; };
int retag_next(const PhaseState &ps, RetagCursor *c, int tid, const char *refname, const char *qn, int start_pos, int hp_of_record);

// parse_variants_for_one_read on the host (only the dropped-interval rescue needs it here;
// the -u path runs the CUDA kernel).  Returns 0 or a POMFRET_GPU_ERR_* code.
struct HostReadVariant { uint32_t pos, len; uint8_t op; uint8_t haptag; };
int host_parse_read_variants(const bam1_t *b, std::vector<HostReadVariant> *out);

}  // namespace pomfret
#endif
