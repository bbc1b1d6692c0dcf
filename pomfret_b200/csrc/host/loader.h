// Host-side read loader: the htslib half of load_reads_given_interval
// (reference blockjoin.c:1043-1138): region query, record filters, HP lookup,
// and packing of the surviving records into pomfret_gpu_read_desc entries.
// Everything after "the record passed the filters" happens on the device.
#ifndef POMFRET_HOST_LOADER_H
#define POMFRET_HOST_LOADER_H
#include <cstdint>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>
#include "htslib/sam.h"
#include "pomfret_gpu.h"

namespace pomfret {

constexpr int kReadback = 50000;        // READBACK, blockjoin.c:19
constexpr int kHaptagUnphased = 254;    // HAPTAG_UNPHASED, blockjoin.c:26
constexpr float kMinAlnDe = 0.1f;       // MIN_ALN_DE, blockjoin.c:23

using RawTagMap = std::unordered_map<std::string, int>;

// One open BAM + index + header (bamfile_t, blockjoin.c:558-593). Opened once per worker.
struct BamReader {
    std::string fn;
    samFile *fp = nullptr;
    hts_idx_t *idx = nullptr;
    sam_hdr_t *hdr = nullptr;
    bam1_t *rec = nullptr;
    bool open(const std::string &path);
    void close();
    ~BamReader() { close(); }
};

// The records of one window that passed the filters, in BAM order, with stable storage.
struct WindowReads {
    uint32_t ref_start = 0, ref_end = 0;
    std::vector<pomfret_gpu_read_desc> descs;
    std::vector<uint32_t> qname_off;  // into qnames, NUL terminated
    std::string qnames;
    std::vector<uint8_t> arena;       // record payload copies; descs point into it
    uint64_t n_bases = 0;
    void clear();
    const char *qname(size_t i) const { return qnames.data() + qname_off[i]; }
};

// get_hp_from_aln, blockjoin.c:910-923
int hp_from_record(const bam1_t *b);

// Fill `desc` from a record. Pointers address b's data.
void describe_record(const bam1_t *b, int hp, pomfret_gpu_read_desc *desc);

// Query chrom:(s-50000)-(e+50000) and hand every record that passes the filters (blockjoin.c:1081-1084) to
// `fn` together with its haplotag; the record is only valid during the call.  Returns 0 or an error code.
int for_each_window_record(BamReader &bam, const char *chrom, uint32_t ref_start, uint32_t ref_end, int readlen_threshold,
                           int min_mapq, const RawTagMap *raw_tags, const std::function<void(const bam1_t *, int hp)> &fn);

// Query chrom:(s-50000)-(e+50000), filter (blockjoin.c:1081-1084) and pack.
// raw_tags: the -u override (blockjoin.c:1114-1122), may be null.
// Returns 0, or a negative POMFRET_GPU_ERR_* code.
int load_window(BamReader &bam, const char *chrom, uint32_t ref_start, uint32_t ref_end, int readlen_threshold,
                int min_mapq, const RawTagMap *raw_tags, WindowReads *out);

}  // namespace pomfret
#endif
