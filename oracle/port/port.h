/* TEST INFRASTRUCTURE — plain-C restatement of the reference's methphase hot
 * path (the "oracle port").  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this; the product never does.
 *
 * Every function cites the reference lines it restates (reference =
 * /root/reference, nanoporetech/pomfret v0.1-r14).  MM/ML semantics come
 * from htslib, an external unpinned dependency that is absent here: they are
 * restated from the SAMtags specification — "parity
 * unpinned" for that layer; everything above it is pinned against the
 * compiled reference (oracle/_ref) by tests/test_oracle_vs_ref.py.
 */
#ifndef POMFRET_ORACLE_PORT_H
#define POMFRET_ORACLE_PORT_H
#include <stdint.h>
#include "../../include/pomfret_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PORT_N_MODS 10 /* reference N_MODS, blockjoin.c:34 */

typedef struct {
    uint32_t *pos;
    uint8_t *cat;
    uint32_t n, m;
} port_calls_t;

/* fill_read_meth_record_from_bam_line + get_mod_poss_on_ref + bam_endpos.
 * Returns the reference's `stat` (1 kept, 0 dropped) or POMFRET_GPU_ERR_FATAL_CIGAR. */
int port_decode_read(const pomfret_gpu_read_desc *r, int lo, int hi, port_calls_t *out, uint32_t *status,
                     uint32_t *end_pos);

/* get_mod_poss_on_ref alone (blockjoin.c:605-792). seq may be NULL (no implicit handling). */
int port_map_mods_to_ref(const uint32_t *cigar, int n_cigar, uint32_t qs, int strand, const uint32_t *mod_pos,
                         const uint8_t *mod_cat, int n_mods, const uint8_t *seq, uint32_t l_qseq,
                         port_calls_t *out);

typedef struct {
    int n;
    uint32_t *real_pos, *starts;
    uint8_t *lens;
} port_sites_t;

typedef struct {
    /* inputs of one window's read set (reads that decoded with stat==1, BAM order) */
    int n;           /* rs->n after the coverage gate */
    int n_loaded;    /* before the gate */
    uint32_t ref_start, ref_end;
    int *hp;         /* current tags (mutated by the greedy loop) */
    uint8_t *strand;
    uint32_t *start_pos, *end_pos;
    port_calls_t *calls;
    uint64_t *revbuf; /* (end<<32|id) ascending */
    uint32_t n_left, n_left_strict, n_right, n_right_strict;
    uint32_t *ids_left, *ids_left_strict, *ids_right, *ids_right_strict;
    /* methmers of the direction currently stored */
    uint32_t **mmr;
    int *mmr_n;
    uint32_t *mmr_start_i;
} port_readset_t;

typedef struct {
    int decision, join_fwd, join_bwd;
    int n_reads, n_reads_loaded;
    int n_sites_fwd, n_sites_bwd;
    int table_fwd[4], table_bwd[4];
    float score_fwd, score_bwd;
    int which_way_fwd, which_way_bwd;
    port_sites_t sites[2];
    uint8_t *tags_final, *tags_fwd, *tags_bwd; /* n_reads_loaded entries */
    int32_t *read_ids;                         /* per input record: id in read set or -1 */
    uint32_t *status;                          /* per input record */
    port_readset_t *rs;                        /* kept for inspection; methmers = fwd direction */
    uint32_t **mmr_bwd; int *mmr_n_bwd; uint32_t *mmr_start_bwd;
    uint32_t *order_fwd, *order_bwd; int n_order_fwd, n_order_bwd; /* greedy tagging order */
    int status_code;
    uint8_t *prop_fwd, *prop_bwd; /* tags right after the greedy loop of each direction */
} port_window_t;

/* The whole of haplotag_region_given_bam (blockjoin.c:4217-4335) on records that already
 * passed the loader's filters. */
port_window_t *port_window_run(const pomfret_gpu_read_desc *reads, int n_reads, uint32_t ref_start, uint32_t ref_end,
                               const pomfret_gpu_config *cfg);
void port_window_free(port_window_t *w);

/* get_methmer_sites_and_ranges (blockjoin.c:3202-3354) */
void port_sites(const port_readset_t *rs, int cov_for_selection, int k, int k_span, int direction, port_sites_t *out);
void port_sites_free(port_sites_t *s);
/* get_mmr_of_read (blockjoin.c:3357-3451); returns number of methmers written to out (malloc'd) */
int port_mmr_of_read(const port_calls_t *calls, const port_sites_t *ms, uint32_t **out, uint32_t *start_i);
/* evaluate_separation1 (blockjoin.c:3881-3939) with the Fisher test; table = buf[ref][query] */
float port_evaluate_separation(const uint8_t *ref, const uint8_t *query, int n, int *join_dir, int table[4]);
double port_fisher_two_sided(int n11, int n12, int n21, int n22);

/* parse_variants_for_one_read + haptag_one_read_with_variants (blockjoin.c:1545-1840).
 * Returns tag {0,1,254} or a negative error code. */
int port_haptag_read(const pomfret_gpu_read_desc *r, const pomfret_gpu_variant *known, uint32_t n_known,
                     const uint8_t *bases, uint32_t known_first, int *votes /* [2] or NULL */);
/* the i_left cursor (blockjoin.c:1716-1720) for a start-sorted list of reads */
void port_haptag_cursors(const uint32_t *start_pos, int n_reads, const pomfret_gpu_variant *known, uint32_t n_known,
                         uint32_t *known_first);

/* merge_close_intervals / lift_decisions / flips / generate_new_phase_blocks (blockjoin.c:2190-2361) live in
 * the host front end's own tests; the oracle for them is oracle/_ref (refh_intervals_*). */

#ifdef __cplusplus
}
#endif
#endif
