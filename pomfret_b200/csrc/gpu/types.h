// Device-side data layout of one batch (DESIGN.md "Data layout in HBM").
#ifndef POMFRET_GPU_TYPES_H
#define POMFRET_GPU_TYPES_H
#include <stdint.h>

namespace pomfret_gpu {

// One staged alignment record, 64 bytes.  Field payloads live in the batch blob at 16-byte
// aligned offsets (stored in units of 16 bytes) so that every warp can stream them with
// 128-bit loads.
struct ReadRec {
    uint32_t pos;        // core.pos
    uint32_t l_qseq;
    uint32_t n_cigar;
    uint32_t flags;      // bits 0-15 BAM flag; 16 tags_malformed; 17 has MM; 18 has ML; 19 has MD
    int32_t hp;
    int32_t mn;          // MN tag or -1
    uint32_t cigar_off, seq_off, mm_off, ml_off, md_off;  // blob offsets / 16
    uint32_t mm_len, ml_len, md_len;
    uint32_t calls_off;  // first slot of this read in the call arrays
    uint32_t calls_cap;  // slots reserved
};
static_assert(sizeof(ReadRec) == 64, "ReadRec must stay 64 bytes");

constexpr uint32_t RF_MALFORMED = 1u << 16;
constexpr uint32_t RF_HAS_MM = 1u << 17;
constexpr uint32_t RF_HAS_ML = 1u << 18;
constexpr uint32_t RF_HAS_MD = 1u << 19;

struct WindowRec {
    uint32_t ref_start, ref_end;
    uint32_t first_read, n_reads;  // candidate records of the window, BAM order
    uint32_t site_off, site_cap;   // slice of the site arrays
    uint32_t calls_begin, calls_end;  // slice of the call arrays spanned by the window's records
};

// Per-window state produced by the read-set / pileup / join kernels.
struct WindowState {
    uint32_t n;             // rs->n after the left-coverage gate
    uint32_t n_loaded;      // reads kept by decode
    uint32_t n_left, n_left_strict, n_right, n_right_strict;
    uint32_t n_sites;
    uint32_t total_calls;   // calls of kept reads
    int32_t status;         // 0 or negative error
    uint32_t mmr_total[2];  // methmers per direction
    uint32_t mmr_base[2];   // offset of the window's slice in the methmer pool
    uint32_t tab_base[2];   // offset (in sites) of the window's count tables
    int32_t table[2][4];    // evaluate_separation1 2x2 tables, [direction][ref*2+query]
    uint32_t n_order[2];
    uint32_t max_cov;       // upper bound of the reads that touch one methmer site (bounds every count of the join tables)
};

// status bits of a decoded read (mirror POMFRET_GPU_READ_* in pomfret_gpu.h)
constexpr uint32_t RS_KEPT = 1u, RS_HAS_IMPLICIT = 2u, RS_FATAL_CIGAR = 4u, RS_MM_ERROR = 8u, RS_SLOWPATH = 16u;
constexpr uint32_t RS_UNSORTED = 32u;  // calls are not strictly increasing (internal)
constexpr uint32_t RS_OVERFLOW = 64u;  // call slots exhausted (internal: engine retries with more room)
constexpr uint32_t RS_LEAN = 128u;     // decoded by the lean path (whole-record tables in shared memory)

constexpr int kMaxK = 8;               // dense count tables (3^k + 1 entries per site) cover methmer keys of up to kMaxK symbols

}  // namespace pomfret_gpu
#endif
