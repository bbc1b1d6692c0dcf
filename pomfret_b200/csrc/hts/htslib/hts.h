/* hts-shim: a small zlib-only implementation of the htslib API subset that
 * Pomfret's methphase path uses (SURVEY.md §8(c) symbol list).
 *
 * htslib itself is an external, un-vendored, unpinned dependency of the
 * reference (reference Makefile:7,11) and is absent from this image, so this
 * shim restates the published formats: BGZF / BAM / BAI from the SAM/BAM
 * specification (SAMv1), MM/ML from the SAMtags specification.
 * "Parity unpinned" against a real libhts (no libhts available to diff with).
 *
 * Declarations only mirror names/field names that callers touch; the layout
 * is this shim's own.
 */
#ifndef POMFRET_HTS_SHIM_HTS_H
#define POMFRET_HTS_SHIM_HTS_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;
#define HTS_POS_MAX ((((int64_t)INT32_MAX) << 32) | INT32_MAX)

struct BGZF;
struct sam_hdr_t;

typedef struct htsFile {
    uint32_t is_bin : 1, is_write : 1, is_be : 1, is_cram : 1, is_bgzf : 1, dummy : 27;
    int64_t lineno;
    char *fn;
    union {
        struct BGZF *bgzf;
        void *voidp;
    } fp;
} htsFile;

typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;

htsFile *hts_open(const char *fn, const char *mode);
int hts_close(htsFile *fp);
void hts_idx_destroy(hts_idx_t *idx);
void hts_itr_destroy(hts_itr_t *itr);
/* shim extension (not htslib API): the (begin, end) virtual-offset pairs of the file chunks of a region iterator */
int pomfret_itr_chunks(const hts_itr_t *itr, const uint64_t **pairs);
int pomfret_idx_linear(const hts_idx_t *idx, int tid, const uint64_t **lin);
int pomfret_idx_nref(const hts_idx_t *idx);

/* shim extension: number of records / uncompressed bytes pulled through
 * iterators by this process (used by the throughput harness). */
void hts_shim_counters(uint64_t *n_records, uint64_t *n_bytes);

#ifdef __cplusplus
}
#endif
#endif
