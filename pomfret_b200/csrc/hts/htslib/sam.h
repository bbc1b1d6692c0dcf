/* hts-shim SAM/BAM record layer: API subset used by Pomfret's methphase path
 * (reference call sites: blockjoin.c:566-593, 794-949, 1043-1173, 1545-1691,
 * 1841-1898, 3022-3103). Formats follow SAMv1 §4.2 (BAM) and §5.2 (BAI). */
#ifndef POMFRET_HTS_SHIM_SAM_H
#define POMFRET_HTS_SHIM_SAM_H
#include <stdint.h>
#include "hts.h"
#include "bgzf.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sam_hdr_t {
    int32_t n_targets;
    size_t l_text;
    uint32_t *target_len;
    char **target_name;
    char *text;
} sam_hdr_t;
typedef sam_hdr_t bam_hdr_t;

#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9

#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define BAM_CIGAR_TYPE 0x3C1A7
#define bam_cigar_op(c) ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_gen(l, o) ((l) << BAM_CIGAR_SHIFT | (o))
/* bit 0: consumes query; bit 1: consumes reference */
#define bam_cigar_type(o) (BAM_CIGAR_TYPE >> ((o) << 1) & 3)

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;
    int l_data;
    uint32_t m_data;
} bam1_t;

/* shim extension (not htslib API): with b->id set to this value, b->data / b->m_data describe a buffer owned by
 * the caller (a loader's record arena); bam_read1 fills it in place and fails instead of reallocating it */
#define POMFRET_BAM_EXTERNAL_DATA 0x504f4d4645584255ull

#define bam_is_rev(b) (((b)->core.flag & BAM_FREVERSE) != 0)
#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) \
    ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b)                                                     \
    ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname +            \
     (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b)                                                   \
    ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname -          \
     (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

extern const char seq_nt16_str[];
extern const unsigned char seq_nt16_table[256];

typedef htsFile samFile;

sam_hdr_t *sam_hdr_read(samFile *fp);
sam_hdr_t *bam_hdr_read(BGZF *fp);
int bam_hdr_write(BGZF *fp, const sam_hdr_t *h);
void sam_hdr_destroy(sam_hdr_t *h);
#define bam_hdr_destroy(h) sam_hdr_destroy(h)
int sam_hdr_name2tid(sam_hdr_t *h, const char *ref);
#define sam_close(fp) hts_close(fp)

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
int bam_read1(BGZF *fp, bam1_t *b);
int bam_write1(BGZF *fp, const bam1_t *b);
hts_pos_t bam_endpos(const bam1_t *b);
int64_t bam_cigar2qlen(int n_cigar, const uint32_t *cigar);
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar);

hts_idx_t *sam_index_load(htsFile *fp, const char *fn);
int sam_index_build3(const char *fn, const char *fnidx, int min_shift, int nthreads);
/* shim extension (not htslib API): BAI construction from a writer's own virtual offsets, see index.c */
typedef struct pomfret_bai_builder pomfret_bai_builder;
pomfret_bai_builder *pomfret_bai_new(int n_ref);
int pomfret_bai_add(pomfret_bai_builder *b, int tid, hts_pos_t beg, hts_pos_t end, int unmapped, uint64_t off0, uint64_t off1);
int pomfret_bai_finish(pomfret_bai_builder *b, const char *fnidx);
hts_itr_t *sam_itr_querys(const hts_idx_t *idx, sam_hdr_t *hdr, const char *region);
hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end);
int sam_itr_next(htsFile *htsfp, hts_itr_t *itr, bam1_t *r);

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int64_t bam_aux2i(const uint8_t *s);
double bam_aux2f(const uint8_t *s);
char *bam_aux2Z(const uint8_t *s);
int bam_aux_update_int(bam1_t *b, const char tag[2], int64_t val);
int bam_aux_append(bam1_t *b, const char tag[2], char type, int len, const uint8_t *data);

/* ---- base modifications (SAMtags MM/ML) ---- */
typedef struct hts_base_mod {
    int modified_base;
    int canonical_base;
    int strand;
    int qual;
} hts_base_mod;
#define HTS_MOD_UNKNOWN -1
#define HTS_MOD_UNCHECKED -2
typedef struct hts_base_mod_state hts_base_mod_state;

hts_base_mod_state *hts_base_mod_state_alloc(void);
void hts_base_mod_state_free(hts_base_mod_state *state);
int bam_parse_basemod(const bam1_t *b, hts_base_mod_state *state);
int bam_mods_at_next_pos(const bam1_t *b, hts_base_mod_state *state, hts_base_mod *mods, int n_mods);
int bam_next_basemod(const bam1_t *b, hts_base_mod_state *state, hts_base_mod *mods, int n_mods, int *pos);

#ifdef __cplusplus
}
#endif
#endif
