/* pomfret_gpu.h — C ABI of libpomfret_gpu, the B200 (sm_100a) engine behind
 * `pomfret methphase` / `pomfret report`.
 *
 * The reference (nanoporetech/pomfret v0.1-r14) has no plugin / FFI layer; the
 * two internal seams this library replaces are:
 *
 *   window engine    dataset_t *haplotag_region_given_bam(storage_t*, char *fn_bam,
 *                        char *chrom, uint32_t ref_start, uint32_t ref_end, mmr_config_t,
 *                        int n_candidates_per_iter, int do_n_permuations, int *decision)
 *                    reference blockjoin.c:4217-4335, called from :4400 and :5058
 *   read haplotagger void pre_haplotagging_read_in_one_ref(char *fn_bam, char *ref_itvl,
 *                        vvar_t *known_vars, htstri_t *qname2haptag_raw)
 *                    reference blockjoin.c:1841-1898, called from :2075 and :2152
 *
 * Two ways in.  (a) Records the caller has already read (htslib side of the seam:
 * BAM iteration and the record filters on the host) are staged with add_reads().
 * (a') Compressed ingest: the caller ships the BGZF blocks of its region queries as
 * they lie in the file; inflate, record walk, filters and field slicing run on the
 * device.  Either way everything from "a record passed the filters" to "decision +
 * one haplotag per read" runs on the device.
 *
 * Conventions
 *   - every entry point returns 0 (POMFRET_GPU_OK) or a negative error code;
 *     the library never calls exit()/abort(): the reference's fatal cases
 *     (e.g. unknown CIGAR op, blockjoin.c:776-778) surface as error codes /
 *     per-read status bits that the host front end turns into the same message;
 *   - the caller owns every buffer it passes in and every output buffer;
 *     add_read() copies, so inputs may be reused as soon as it returns;
 *   - the library owns device memory and pinned staging arenas;
 *   - a ctx is thread-safe; a batch is used by one thread at a time and is bound
 *     to one device and one stream; results are valid when collect() returns;
 *   - there is no CPU fallback: init fails if no CUDA device is usable.
 */
#ifndef POMFRET_GPU_H
#define POMFRET_GPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POMFRET_GPU_OK 0
#define POMFRET_GPU_ERR_NO_DEVICE -1
#define POMFRET_GPU_ERR_CUDA -2
#define POMFRET_GPU_ERR_NOMEM -3
#define POMFRET_GPU_ERR_ARG -4
#define POMFRET_GPU_ERR_STATE -5
#define POMFRET_GPU_ERR_FATAL_CIGAR -6  /* reference: "fatal: unknown cigar operation" blockjoin.c:777 */
#define POMFRET_GPU_ERR_DUP_QNAME -7    /* reserved for the host loader, blockjoin.c:1149 */
#define POMFRET_GPU_ERR_UNSUPPORTED -8  /* input outside the engine's limits: methmer k > 8 (dense 3^k count tables), more than 1024
                                           candidates per iteration, 60000 or more records in one window (16-bit read ids),
                                           a read of 2^28 bases or more */
#define POMFRET_GPU_ERR_MISSING_MD -9   /* reference: assert(tagd) blockjoin.c:1596 */
#define POMFRET_GPU_ERR_BAD_MD -10      /* reference: "invalid MD" blockjoin.c:1622 */

typedef struct pomfret_gpu_ctx pomfret_gpu_ctx;
typedef struct pomfret_gpu_batch pomfret_gpu_batch;

/* Mirrors mmr_config_t (reference blockjoin.h:7-16) plus the per-call scalars of
 * haplotag_region_given_bam. */
typedef struct pomfret_gpu_config {
    int32_t k;
    int32_t k_span;
    int32_t lo, hi;
    int32_t cov_known;
    int32_t cov_for_selection;
    int32_t cov_for_runtime;
    int32_t readlen_threshold;
    int32_t min_mapq;
    int32_t n_candidates_per_iter;
} pomfret_gpu_config;

/* One alignment record that passed the host-side filters of
 * load_reads_given_interval (blockjoin.c:1081-1084).  Pointers address the
 * caller's copy of the BAM record. */
typedef struct pomfret_gpu_read_desc {
    uint32_t pos;         /* core.pos */
    uint32_t l_qseq;      /* core.l_qseq */
    uint32_t n_cigar;     /* core.n_cigar */
    uint16_t flag;        /* core.flag */
    uint8_t mapq;         /* core.qual */
    uint8_t tags_malformed; /* 1: MM not 'Z' / ML not 'B,C' / MN mismatch handled by caller => record has no mods */
    int32_t hp;           /* get_hp_from_aln (blockjoin.c:910-923) or the qname2haptag_raw override (:1114-1122) */
    int32_t mn;           /* value of the MN tag, -1 if absent */
    const uint32_t *cigar;
    const uint8_t *seq;   /* 4-bit packed, (l_qseq+1)/2 bytes */
    const char *mm;       /* MM/Mm string without the type byte, not necessarily NUL terminated; NULL if absent */
    uint32_t mm_len;
    int32_t ml_len;       /* number of ML bytes, -1 if the tag is absent */
    const uint8_t *ml;
    const char *md;       /* MD string (only the haplotagger needs it); NULL if absent */
    uint32_t md_len;
    uint32_t reserved;
} pomfret_gpu_read_desc;

/* Phased variant of the known set, as produced by insert_variant_from_vcf_line
 * (blockjoin.c:1432-1543): pos is 0-based (deletions: +1), op 1=X 2=I 3=D,
 * bases in seq_nt4 code (A0 C1 G2 T3 other 4), haptag = haplotype of the REF allele. */
typedef struct pomfret_gpu_variant {
    uint32_t pos;
    uint32_t len;
    uint8_t op;
    uint8_t haptag;
    uint16_t reserved;
    uint32_t bases_off;   /* offset of the first base in the `bases` array given to haptag() */
} pomfret_gpu_variant;

typedef struct pomfret_gpu_window_result {
    int32_t decision;      /* 0 cis, 1 trans, -1 no join (blockjoin.c:4313-4320) */
    int32_t join_fwd;      /* join1, haplotag_region2 direction 0 */
    int32_t join_bwd;      /* join2, direction 1 */
    int32_t n_reads;       /* rs->n after the left-coverage gate (0 if abandoned, blockjoin.c:1161-1163) */
    int32_t n_reads_loaded;/* reads that produced a record before the gate */
    int32_t n_sites_fwd, n_sites_bwd;
    int32_t n_left, n_left_strict, n_right, n_right_strict;
    int32_t table_fwd[4];  /* evaluate_separation1 2x2 table buf[ref][query], direction 0 (right strict reads) */
    int32_t table_bwd[4];  /* same for direction 1 (left strict reads) */
    float score_fwd, score_bwd;
    int32_t which_way_fwd, which_way_bwd;
    int32_t status;        /* 0, or a negative error code raised inside this window */
} pomfret_gpu_window_result;

/* per-read status bits written by decode() */
#define POMFRET_GPU_READ_KEPT 1u          /* add_read_record_from_bam_line returned 1 */
#define POMFRET_GPU_READ_HAS_IMPLICIT 2u  /* has_implicit, blockjoin.c:856 */
#define POMFRET_GPU_READ_FATAL_CIGAR 4u
#define POMFRET_GPU_READ_MM_ERROR 8u      /* malformed MM/ML: record has no modifications */
#define POMFRET_GPU_READ_SLOWPATH 16u     /* decoded by the single-lane general path */
#define POMFRET_GPU_READ_UNSORTED 32u     /* internal: calls not strictly ascending */
#define POMFRET_GPU_READ_LEAN 128u        /* decoded by the lean path (whole-record tables in shared memory) */
#define POMFRET_GPU_READ_OVERFLOW 64u     /* internal: ran out of call slots; pileup() re-runs the record with more room */

/* ---- context ---- */
int pomfret_gpu_init(pomfret_gpu_ctx **out, const int *devices, int n_devices, int n_workers);
void pomfret_gpu_destroy(pomfret_gpu_ctx *ctx);
const char *pomfret_gpu_strerror(int rc);
int pomfret_gpu_device_count(void);
const char *pomfret_gpu_version(void);

/* ---- (a) staging ---- */
/* Optional: pin and map a host buffer that holds alignment records (e.g. the buffer BGZF blocks are inflated
 * into) for as long as it is registered.  add_reads() then leaves the payload of records that lie completely
 * inside registered buffers where it is, and submit() lets the device gather it over PCIe: no host-side copy.
 * Records elsewhere are copied into the library's pinned arena as before.  The buffer must not be freed,
 * moved or modified between add_reads() and the completion of submit()'s transfers (collect(), rewind() or
 * reset() of the batch). */
int pomfret_gpu_host_register(pomfret_gpu_ctx *ctx, void *ptr, size_t bytes);
int pomfret_gpu_host_unregister(pomfret_gpu_ctx *ctx, void *ptr);
int pomfret_gpu_batch_begin(pomfret_gpu_ctx *ctx, int worker, int device, pomfret_gpu_batch **out);
int pomfret_gpu_batch_reset(pomfret_gpu_batch *b);
int pomfret_gpu_batch_add_read(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r);
/* the same for an array of n descriptors (one call per window instead of one per record) */
int pomfret_gpu_batch_add_reads(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r, uint32_t n);
/* ---- (a') compressed ingest: BGZF inflate + BAM record slicing on the device ----
 * Replaces, for the engine's inputs, what the reference does through htslib per window: sam_itr_next / bam_read1 /
 * bgzf_read at blockjoin.c:1070 (windows), :1863 (-u pre-pass); the record filters of :1081-1084 run on the device too.
 * The caller ships the BGZF blocks of the index chunks of its region queries exactly as they lie in the file and
 * never inflates or parses a record. */
typedef struct pomfret_gpu_bgzf_block {
    uint64_t comp_off;   /* offset of the block (its gzip header) in the compressed buffer */
    uint32_t csize;      /* whole block: BSIZE + 1 */
    uint32_t isize;      /* ISIZE of the block's footer */
    uint64_t out_off;    /* where the block inflates to: the blocks of a stream back to back */
} pomfret_gpu_bgzf_block;
typedef struct pomfret_gpu_bgzf_stream {  /* one chunk of the index: consecutive blocks that hold whole records */
    uint64_t out_off;    /* out_off of the stream's first block */
    uint64_t out_bytes;  /* inflated bytes up to the end of the chunk's last record */
    uint32_t ubeg;       /* offset of the chunk's first record inside the first block */
    int32_t tid;         /* a record of another target ends the stream ... */
    uint32_t end0;       /* ... and so does one that starts at or behind this position (the query's end, 0-based exclusive) */
    uint32_t first_block, n_blocks;
    uint32_t reserved;
} pomfret_gpu_bgzf_stream;
typedef struct pomfret_gpu_ingest_filter {  /* blockjoin.c:1081-1084; all zero = primary records only (blockjoin.c:1862) */
    uint32_t min_mapq;
    uint32_t min_len;        /* readlen_threshold */
    uint32_t min_len_floor;  /* 2: "len < 2" of the window loader */
    uint32_t check_de;
    float max_de;            /* MIN_ALN_DE 0.1 */
    uint32_t keep_all_flags; /* 1: unmapped / secondary / supplementary records pass too (blockjoin.c:2512: no flag test) */
} pomfret_gpu_ingest_filter;
typedef struct pomfret_gpu_sliced_record {  /* one alignment record as the slicing kernel saw it; addresses are DEVICE addresses */
    uint32_t pos, end_pos, l_qseq, n_cigar;   /* n_cigar / cigar: the real operations, also for CG-tag records */
    uint16_t flag;
    uint8_t mapq, tags_malformed;
    int32_t hp, mn;                           /* as in pomfret_gpu_read_desc (hp: 254 if absent) */
    uint32_t mm_len;
    int32_t ml_len;
    uint32_t md_len;
    uint32_t stream;
    uint8_t keep;                             /* passed the filters */
    uint8_t bad;                              /* malformed record (fields run past its end) */
    uint8_t has_mm, hp_irregular;             /* MM/Mm present; HP:0 ("irregular HP tag", blockjoin.c:916) */
    uint8_t l_qname;                          /* with the NUL */
    uint8_t hp_type;                          /* type byte of the first HP tag (0 if the record has none) */
    uint8_t cg_cigar;                         /* 1: the CIGAR was taken from the CG tag */
    uint8_t pad;
    uint64_t cigar, seq, mm, ml, md;          /* 0 if absent */
    uint64_t qname_dev;
    uint32_t rec_bytes;                       /* the record in the file: 4 + block_size */
    uint32_t hp_off;                          /* offset of the HP tag's type byte from the record's start (if hp_type) */
    int32_t tid;
    uint32_t reserved;
    char qname[48];                           /* NUL terminated prefix; the whole name if l_qname <= 48 */
} pomfret_gpu_sliced_record;
#define POMFRET_GPU_ANY_TID (-0x7fffffff - 1)  /* pomfret_gpu_bgzf_stream::tid: walk every record of the stream, whatever its target or position */
/* pinned host buffer of the batch for the compressed blocks (valid until the next reset): the loader reads the file into it */
int pomfret_gpu_batch_ingest_buffer(pomfret_gpu_batch *b, size_t bytes, void **out);
/* H2D of the blocks, inflate (ISIZE + CRC-32 checked), record walk and slicing; *n_records = records found in the streams,
 * in stream order and file order inside a stream.  `comp` is the buffer of ingest_buffer() or any host memory. */
int pomfret_gpu_batch_ingest_bgzf(pomfret_gpu_batch *b, const void *comp, size_t comp_bytes, const pomfret_gpu_bgzf_block *blocks,
                                  uint32_t n_blocks, const pomfret_gpu_bgzf_stream *streams, uint32_t n_streams,
                                  const pomfret_gpu_ingest_filter *flt, uint32_t *n_records);
int pomfret_gpu_batch_ingest_records(pomfret_gpu_batch *b, pomfret_gpu_sliced_record *out, uint32_t cap);
/* Coverage estimate over the records of the ingest (estimate_read_coverage_dirtyfast, blockjoin.c:951-1040; SURVEY.md
 * §8(f) row 4): with the filter set to {min_mapq 5, min_len 15000, check_de} the kept records that start at or behind
 * min_pos each add one to bin i/bin_size for i = pos, pos+bin_size, ... < end_pos; *increments = how many of those
 * fall on a bin below n_bins (the reference only uses the sum of a contig's bins: coverage = sum / n_bins). */
int pomfret_gpu_batch_ingest_coverage(pomfret_gpu_batch *b, uint32_t min_pos, uint32_t bin_size, uint32_t n_bins, uint64_t *increments);
/* Output-BAM re-tagging (output_modify_bam, blockjoin.c:3022-3103).  Every record of the ingest, in
 * order, is copied into one contiguous uncompressed BAM stream with its HP tag set the way bam_aux_update_int(aln, "HP", v)
 * does it (blockjoin.c:3092): hp_val[i] = v in 1..255 (values below 255 as type C, 255 as S); an integer HP tag of at least
 * that size is overwritten in place (keeping its size, unsigned type), a smaller one grows, a record without the tag gets
 * it appended; hp_val[i] = 0 copies the record unchanged.  dst_off[i] = offset of record i in the output stream (the caller
 * derives the new sizes from rec_bytes / hp_type of the sliced records); the out_bytes bytes of the stream are written to
 * host_out. */
int pomfret_gpu_batch_ingest_retag(pomfret_gpu_batch *b, const uint64_t *dst_off, const uint8_t *hp_val, uint64_t out_bytes, void *host_out);
/* add_reads_shared() for records of the ingest: the pointers of r[] are the device addresses of a sliced record and
 * r[i].reserved carries its end_pos; the payload is copied out of the inflated stream on the device. */
int pomfret_gpu_batch_add_reads_device(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r, uint32_t n, const int64_t *same_as);
/* parity getters: the inflated bytes of one stream; a record's whole name */
int pomfret_gpu_debug_get_inflated(pomfret_gpu_batch *b, uint32_t stream, uint8_t *dst, size_t cap, size_t *n);
int pomfret_gpu_batch_ingest_qname(pomfret_gpu_batch *b, uint32_t record, char *dst, uint32_t cap);

/* Decode-once form: same_as[i] >= 0 names an earlier read of this batch that is the same
 * alignment record (it lies in two overlapping windows; the reference re-decodes it per window, blockjoin.c:1056).
 * The new slot shares that read's staged payload, call slots and decode result; r[i] then only needs pos, l_qseq,
 * n_cigar (checked) and hp.  same_as[i] < 0 (or same_as == NULL): an ordinary record. */
int pomfret_gpu_batch_add_reads_shared(pomfret_gpu_batch *b, const pomfret_gpu_read_desc *r, uint32_t n, const int64_t *same_as);
/* reads [first_read, first_read+n_reads) are the records of the region query
 * chrom:(ref_start-50000)-(ref_end+50000) in BAM order */
int pomfret_gpu_batch_add_window(pomfret_gpu_batch *b, uint32_t ref_start, uint32_t ref_end, uint32_t first_read,
                                 uint32_t n_reads);
/* the same for n windows at once (parallel arrays) */
int pomfret_gpu_batch_add_windows(pomfret_gpu_batch *b, const uint32_t *ref_start, const uint32_t *ref_end,
                                  const uint32_t *first_read, const uint32_t *n_reads, uint32_t n);
int pomfret_gpu_batch_submit(pomfret_gpu_batch *b); /* async H2D of the staged records */
/* Back to the state right after submit(): the staged records stay resident on the device and the stages can
 * be run again (other thresholds, k, candidate counts on the same reads).  Waits for the batch's stream. */
int pomfret_gpu_batch_rewind(pomfret_gpu_batch *b);

/* ---- (b) MM/ML + CIGAR decode: fill_read_meth_record_from_bam_line + get_mod_poss_on_ref ---- */
int pomfret_gpu_decode(pomfret_gpu_batch *b, uint8_t qual_lo, uint8_t qual_hi);
/* ---- (c) read haplotagging: parse_variants_for_one_read + haptag_one_read_with_variants.
 * known_first[i] is the i_left cursor of read i (blockjoin.c:1716-1720), computed by the caller. */
int pomfret_gpu_haptag(pomfret_gpu_batch *b, const pomfret_gpu_variant *known, uint32_t n_known,
                       const uint8_t *bases, uint32_t n_bases, const uint32_t *known_first);
/* ---- (c') phase of unphased variants inside dropped intervals: recover_variant_phase_in_one_interval, blockjoin.c:2475-2600
 * (SURVEY.md §8(f) row 3).  After haptag() — which parses the variants every record shows itself, with or without a
 * known set — count, per known position (ascending), the records with read_hap 0 / 1 that show a variant there:
 * votes[2*i + hap].  read_hap[i] = 255 leaves record i out.  votes[2*n_positions] = read variants at or behind the last
 * known position (the reference never evaluates a known variant that ends its merged list, :2561). */
int pomfret_gpu_variant_votes(pomfret_gpu_batch *b, const uint32_t *positions, uint32_t n_positions, const uint8_t *read_hap,
                              int32_t *votes /* 2 * n_positions + 1 */);
/* ---- (d) read set, CpG site pileup, methmer layout + extraction ---- */
int pomfret_gpu_pileup(pomfret_gpu_batch *b, const pomfret_gpu_config *cfg);
/* ---- (e) greedy propagation in both directions + join evaluation ---- */
int pomfret_gpu_join(pomfret_gpu_batch *b, const pomfret_gpu_config *cfg);

/* Synchronise and fetch: one result per window; read_tags[i] = final hp of batch read i in its
 * window (rs->a[j].hp after haplotag_region_given_bam; 255 for records that were not kept);
 * read_ids[i] = index j of the read inside its window's read set, -1 if not kept.  Either
 * output pointer may be NULL. */
int pomfret_gpu_batch_collect(pomfret_gpu_batch *b, pomfret_gpu_window_result *win, uint8_t *read_tags,
                              int32_t *read_ids);
/* haplotagger output: tags[i] in {0,1,254}; status[i] 0 or a negative code */
int pomfret_gpu_batch_collect_haptags(pomfret_gpu_batch *b, uint8_t *tags, int32_t *status);
void pomfret_gpu_batch_end(pomfret_gpu_batch *b);

/* ---- timing of the last run of each stage on this batch (CUDA events, ms) ---- */
typedef struct pomfret_gpu_timing {
    float h2d_ms, decode_ms, haptag_ms, readset_ms, pileup_ms, methmer_ms, join_ms, d2h_ms;
    uint64_t bytes_h2d, bytes_d2h;
    uint64_t decode_bytes, pileup_bytes, methmer_bytes, haptag_bytes; /* algorithmic bytes, SURVEY.md §8(d) */
    uint32_t launches;
    float inflate_ms, slice_ms;               /* compressed ingest: inflate_kernel; walk + scan + slice kernels */
    uint64_t inflate_in_bytes, inflate_out_bytes;
} pomfret_gpu_timing;
int pomfret_gpu_batch_timing(pomfret_gpu_batch *b, pomfret_gpu_timing *out);

/* ---- parity / debug getters (valid after the stage that produces them; they synchronise) ---- */
int pomfret_gpu_debug_read_info(pomfret_gpu_batch *b, uint32_t read, uint32_t *status, uint32_t *n_calls,
                                uint32_t *end_pos);
int pomfret_gpu_debug_get_calls(pomfret_gpu_batch *b, uint32_t read, uint32_t *pos, uint8_t *cat, uint32_t cap,
                                uint32_t *n);
int pomfret_gpu_debug_get_sites(pomfret_gpu_batch *b, uint32_t window, int direction, uint32_t *real_pos,
                                uint32_t *starts, uint8_t *lens, uint32_t cap, uint32_t *n);
int pomfret_gpu_debug_get_mmrs(pomfret_gpu_batch *b, uint32_t read, int direction, uint32_t *mmr, uint32_t cap,
                               uint32_t *n, uint32_t *start_i);
int pomfret_gpu_debug_get_tags(pomfret_gpu_batch *b, int direction, uint8_t *tags /* one per batch read */);
/* tagging order of the greedy loop: read ids (within window) in the order they were tagged */
int pomfret_gpu_debug_get_tag_order(pomfret_gpu_batch *b, uint32_t window, int direction, uint32_t *ids,
                                    uint32_t cap, uint32_t *n);
/* hp_cnt[0], hp_cnt[1] of haptag_one_read_with_variants (blockjoin.c:1746), two per read */
int pomfret_gpu_debug_get_votes(pomfret_gpu_batch *b, int32_t *votes);

#ifdef __cplusplus
}
#endif
#endif
