/* Synthetic methphase input generator; see synth.cpp. */
#ifndef POMFRET_SYNTH_H
#define POMFRET_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct synth_config {
    uint64_t seed;
    double coverage;        /* total depth (both haplotypes) */
    double read_len_mean;   /* lognormal mean, reference bases */
    double read_len_sigma;  /* lognormal sigma */
    int64_t read_len_min;
    int cpg_period;         /* one designated CpG per ~cpg_period bp */
    int var_period;         /* one germline variant per ~var_period bp */
    double indel_var_frac;
    double block_len_median, block_len_sigma, short_block_frac;
    int64_t gap_min, gap_max;
    double err_rate;        /* per-base sequencing error (40% X, 30% I, 30% D) */
    double softclip_frac;
    double frac_meth, frac_unmeth; /* rest is haplotype specific */
    double ml_flip, ml_mid;
    double hp_drop;
    int tagged;             /* write HP/PS tags (0 = Dorado-style untagged BAM) */
    int qual_mode;          /* 0 = missing qualities (0xff), 1 = pseudo-random */
    int compress_level;     /* BGZF deflate level 0..9 */
    double frac_low_mapq, frac_high_de, frac_secondary, frac_no_mm;
    double frac_two_segments, frac_multicode, frac_noncpg_calls;
    int n_header_contigs_before; /* filler @SQ lines so target ids are not 0 */
    double frac_cpg_listed;      /* fraction of CpG cytosines that appear in the MM list (1 = all) */
    double de_cap;               /* > 0: upper bound of the reported `de` tag (keeps very noisy reads past the filter) */
    const char *vcf_in;          /* variants, phase sets and the header's contig list come from this VCF (plain or
                                    gzip) instead of being simulated; no <prefix>.vcf.gz is written */
} synth_config;

void pomfret_synth_default_config(synth_config *c);

/* Simulate reads on each contig inside [region_beg[i], region_end[i]) (NULL or
 * region_end<=0: whole contig) and write <prefix>.bam, .bam.bai, .vcf.gz,
 * .truth.tsv.  Returns 0 on success. */
int pomfret_synth_write(const synth_config *cfg, const char *const *contig_names, const int64_t *contig_lens,
                        const int64_t *region_beg, const int64_t *region_end, int n_contigs, const char *prefix);

#ifdef __cplusplus
}
#endif
#endif
