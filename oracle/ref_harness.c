/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Harness around the UNMODIFIED reference implementation: this translation
 * unit #includes /root/reference/blockjoin.c where it lies (nothing is copied
 * into the repo) and exposes flat-array entry points for the tests.  It is
 * built by oracle/Makefile into oracle/_ref/libpomfret_ref.so together with
 * the reference's cli.c / kstring.c / kthread.c and this repo's hts-shim
 * (htslib is not available in the image).
 *
 * The per-window driver below replays haplotag_region_given_bam
 * (blockjoin.c:4217-4335) call by call so that intermediate state (sites,
 * methmers and the propagated tags of each direction) can be captured;
 * refwin_run_whole() calls the reference function itself and is used to check
 * that the replay is faithful.
 */
#include "blockjoin.c"
#include <unistd.h>
#include <fcntl.h>

typedef struct {
    int n_reads_loaded, n_reads;
    int decision, join1, join2, skipped;
    int n_sites_fwd, n_sites_bwd;
    uint32_t *sites_fwd, *starts_fwd, *sites_bwd, *starts_bwd;
    uint8_t *lens_fwd, *lens_bwd;
    int *hp_init;
    uint8_t *strand;
    uint32_t *len;
    uint32_t *calls_off, *calls_pos;
    uint8_t *calls_cat;
    uint32_t *mmr_off_fwd, *mmr_fwd, *mmr_start_fwd;
    uint32_t *mmr_off_bwd, *mmr_bwd, *mmr_start_bwd;
    uint8_t *tags_fwd, *tags_bwd, *tags_final;
    uint64_t *revbuf;
    uint32_t n_left, n_left_strict, n_right, n_right_strict;
    uint32_t *ids_left, *ids_left_strict, *ids_right, *ids_right_strict;
    char *qnames;
    uint32_t qnames_len;
    uint32_t *qname_off;
} refwin_t;

static int g_saved_stderr = -1;
void refh_quiet(int on) {
    fflush(stderr);
    if (on && g_saved_stderr < 0) {
        g_saved_stderr = dup(2);
        int fd = open("/dev/null", O_WRONLY);
        dup2(fd, 2);
        close(fd);
    } else if (!on && g_saved_stderr >= 0) {
        dup2(g_saved_stderr, 2);
        close(g_saved_stderr);
        g_saved_stderr = -1;
    }
}

static uint32_t *dup_u32(const uint32_t *a, size_t n) {
    uint32_t *r = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    if (n) memcpy(r, a, sizeof(uint32_t) * n);
    return r;
}
static uint8_t *dup_u8(const uint8_t *a, size_t n) {
    uint8_t *r = (uint8_t *)malloc(n ? n : 1);
    if (n) memcpy(r, a, n);
    return r;
}

static void snapshot_mmrs(rs_t *rs, uint32_t **off, uint32_t **mmr, uint32_t **start) {
    size_t tot = 0;
    for (uint32_t i = 0; i < rs->n; i++) tot += rs->a[i].mmr_n;
    *off = (uint32_t *)malloc(sizeof(uint32_t) * (rs->n + 1));
    *mmr = (uint32_t *)malloc(sizeof(uint32_t) * (tot ? tot : 1));
    *start = (uint32_t *)malloc(sizeof(uint32_t) * (rs->n ? rs->n : 1));
    size_t o = 0;
    for (uint32_t i = 0; i < rs->n; i++) {
        (*off)[i] = (uint32_t)o;
        (*start)[i] = rs->a[i].mmr_start_i;
        if (rs->a[i].mmr_n) memcpy(*mmr + o, rs->a[i].mmr, sizeof(uint32_t) * rs->a[i].mmr_n);
        o += rs->a[i].mmr_n;
    }
    (*off)[rs->n] = (uint32_t)o;
}

static uint8_t *snapshot_tags(rs_t *rs) {
    uint8_t *t = (uint8_t *)malloc(rs->n ? rs->n : 1);
    for (uint32_t i = 0; i < rs->n; i++) t[i] = (uint8_t)rs->a[i].hp;
    return t;
}

static mmr_config_t make_cfg(int k, int k_span, int lo, int hi, int cov_known, int cov_sel, int cov_run,
                             int readlen_thr, int min_mapq) {
    mmr_config_t c;
    c.k = k; c.k_span = k_span; c.lo = lo; c.hi = hi;
    c.cov_known = cov_known; c.cov_for_selection = cov_sel; c.cov_for_runtime = cov_run;
    c.readlen_threshold = readlen_thr; c.min_mapq = min_mapq;
    return c;
}

/* raw_tags: optional htstri_t* from refh_pre_haplotag (the -u path) */
refwin_t *refwin_run(const char *fn_bam, const char *chrom, uint32_t ref_start, uint32_t ref_end, int k, int k_span,
                     int lo, int hi, int cov_known, int cov_sel, int cov_run, int readlen_thr, int min_mapq,
                     int n_cand, void *raw_tags) {
    mmr_config_t cfg = make_cfg(k, k_span, lo, hi, cov_known, cov_sel, cov_run, readlen_thr, min_mapq);
    refwin_t *w = (refwin_t *)calloc(1, sizeof(refwin_t));
    kstring_t log = {0, 0, 0};
    rs_t *rs = load_reads_given_interval((char *)fn_bam, (char *)chrom, ref_start, ref_end, READBACK, cfg,
                                         (htstri_t *)raw_tags, &log);
    w->n_reads_loaded = (int)rs->revbuf.n;
    w->n_reads = (int)rs->n;
    w->decision = w->join1 = w->join2 = -1;

    /* read set snapshot (uses revbuf.n so that abandoned windows still show what was loaded) */
    uint32_t nl = (uint32_t)rs->revbuf.n;
    w->hp_init = (int *)malloc(sizeof(int) * (nl ? nl : 1));
    w->strand = (uint8_t *)malloc(nl ? nl : 1);
    w->len = (uint32_t *)malloc(sizeof(uint32_t) * (nl ? nl : 1));
    w->calls_off = (uint32_t *)malloc(sizeof(uint32_t) * (nl + 1));
    w->qname_off = (uint32_t *)malloc(sizeof(uint32_t) * (nl + 1));
    size_t tot = 0;
    for (uint32_t i = 0; i < nl; i++) tot += rs->a[i].meth.calls.n;
    w->calls_pos = (uint32_t *)malloc(sizeof(uint32_t) * (tot ? tot : 1));
    w->calls_cat = (uint8_t *)malloc(tot ? tot : 1);
    size_t o = 0;
    for (uint32_t i = 0; i < nl; i++) {
        w->hp_init[i] = rs->a[i].hp;
        w->strand[i] = rs->a[i].strand;
        w->len[i] = rs->a[i].len;
        w->calls_off[i] = (uint32_t)o;
        memcpy(w->calls_pos + o, rs->a[i].meth.calls.a, sizeof(uint32_t) * rs->a[i].meth.calls.n);
        memcpy(w->calls_cat + o, rs->a[i].meth.quals.a, rs->a[i].meth.quals.n);
        o += rs->a[i].meth.calls.n;
        w->qname_off[i] = rs->names_acl.a[i];
    }
    w->calls_off[nl] = (uint32_t)o;
    w->qname_off[nl] = rs->names_n;
    w->qnames_len = rs->names_n;
    w->qnames = (char *)malloc(rs->names_n ? rs->names_n : 1);
    memcpy(w->qnames, rs->names, rs->names_n);
    w->revbuf = (uint64_t *)malloc(sizeof(uint64_t) * (nl ? nl : 1));
    memcpy(w->revbuf, rs->revbuf.a, sizeof(uint64_t) * nl);
    w->n_left = rs->rf->IDs_left.n; w->ids_left = dup_u32(rs->rf->IDs_left.a, w->n_left);
    w->n_left_strict = rs->rf->IDs_left_strict.n; w->ids_left_strict = dup_u32(rs->rf->IDs_left_strict.a, w->n_left_strict);
    w->n_right = rs->rf->IDs_right.n; w->ids_right = dup_u32(rs->rf->IDs_right.a, w->n_right);
    w->n_right_strict = rs->rf->IDs_right_strict.n; w->ids_right_strict = dup_u32(rs->rf->IDs_right_strict.a, w->n_right_strict);

    /* ---- replay of blockjoin.c:4250-4320 ---- */
    methmers_t *ms = get_methmer_sites_and_ranges(rs, cfg, 0, NULL, NULL);
    methmers_t *ms_bwd = get_methmer_sites_and_ranges(rs, cfg, 1, NULL, NULL);
    w->n_sites_fwd = ms->n;
    w->n_sites_bwd = ms_bwd->n;
    w->sites_fwd = dup_u32(ms->sites_real_poss, ms->n);
    w->starts_fwd = dup_u32(ms->sites_starts, ms->n);
    w->lens_fwd = dup_u8(ms->mmr_lens, ms->n);
    w->sites_bwd = dup_u32(ms_bwd->sites_real_poss, ms_bwd->n);
    w->starts_bwd = dup_u32(ms_bwd->sites_starts, ms_bwd->n);
    w->lens_bwd = dup_u8(ms_bwd->mmr_lens, ms_bwd->n);
    if (ms->n == 0 || ms_bwd->n == 0) {
        w->skipped = 1;
        w->tags_final = snapshot_tags(rs);
        w->tags_fwd = snapshot_tags(rs);
        w->tags_bwd = snapshot_tags(rs);
        snapshot_mmrs(rs, &w->mmr_off_fwd, &w->mmr_fwd, &w->mmr_start_fwd);
        snapshot_mmrs(rs, &w->mmr_off_bwd, &w->mmr_bwd, &w->mmr_start_bwd);
    } else {
        vu32_t readIDs;
        kv_init(readIDs);
        kv_resize(uint32_t, readIDs, 128);
        for (uint32_t i = 0; i < rs->n; i++) kv_push(uint32_t, readIDs, i);
        vu8_t *initial = store_haplotags(rs);

        store_mmr_of_reads(rs, ms_bwd);
        snapshot_mmrs(rs, &w->mmr_off_bwd, &w->mmr_bwd, &w->mmr_start_bwd);
        w->join2 = haplotag_region2(rs, ms_bwd, readIDs.a, readIDs.n, 1, n_cand, cfg.cov_for_runtime, 1, 0, &log);
        w->tags_bwd = snapshot_tags(rs);
        restore_haplotags(rs, initial); /* == do_reset=1 of the reference call */
        wipe_mmr_of_reads(rs);
        store_mmr_of_reads(rs, ms);
        snapshot_mmrs(rs, &w->mmr_off_fwd, &w->mmr_fwd, &w->mmr_start_fwd);
        w->join1 = haplotag_region2(rs, ms, readIDs.a, readIDs.n, 0, n_cand, cfg.cov_for_runtime, 1, 0, &log);
        w->tags_fwd = snapshot_tags(rs);
        if (w->join1 != w->join2 || (w->join1 == -1 && w->join2 == -1)) {
            set_all_as_unphased(rs);
            w->decision = -1;
        } else w->decision = w->join1;
        w->tags_final = snapshot_tags(rs);
        kv_destroy(*initial);
        free(initial);
        kv_destroy(readIDs);
    }
    free(log.s);
    destroy_rs_t(rs);
    destroy_methmers_t(ms);
    destroy_methmers_t(ms_bwd);
    return w;
}

void refwin_free(refwin_t *w) {
    if (!w) return;
    free(w->sites_fwd); free(w->starts_fwd); free(w->sites_bwd); free(w->starts_bwd);
    free(w->lens_fwd); free(w->lens_bwd); free(w->hp_init); free(w->strand); free(w->len);
    free(w->calls_off); free(w->calls_pos); free(w->calls_cat);
    free(w->mmr_off_fwd); free(w->mmr_fwd); free(w->mmr_start_fwd);
    free(w->mmr_off_bwd); free(w->mmr_bwd); free(w->mmr_start_bwd);
    free(w->tags_fwd); free(w->tags_bwd); free(w->tags_final); free(w->revbuf);
    free(w->ids_left); free(w->ids_left_strict); free(w->ids_right); free(w->ids_right_strict);
    free(w->qnames); free(w->qname_off);
    free(w);
}

/* The reference function itself: decision + final tags (tags_out has room for cap entries). */
int refwin_run_whole(const char *fn_bam, const char *chrom, uint32_t ref_start, uint32_t ref_end, int k, int k_span,
                     int lo, int hi, int cov_known, int cov_sel, int cov_run, int readlen_thr, int min_mapq,
                     int n_cand, void *raw_tags, uint8_t *tags_out, int cap, int *n_reads) {
    mmr_config_t cfg = make_cfg(k, k_span, lo, hi, cov_known, cov_sel, cov_run, readlen_thr, min_mapq);
    storage_t st;
    memset(&st, 0, sizeof(st));
    st.stores_raw_tag = raw_tags != NULL;
    st.qname2haptag_raw = (htstri_t *)raw_tags;
    int decision = -1;
    dataset_t *ds = haplotag_region_given_bam(&st, (char *)fn_bam, (char *)chrom, ref_start, ref_end, cfg, n_cand, 1,
                                              &decision);
    *n_reads = (int)ds->rs->n;
    for (int i = 0; i < (int)ds->rs->n && i < cap; i++) tags_out[i] = (uint8_t)ds->rs->a[i].hp;
    destroy_dataset_t(ds, 1);
    return decision;
}

/* ---- direct access to get_mod_poss_on_ref (SURVEY.md App. B.1 vectors) ----
 * Returns the function's return value; *n_out receives the number of calls. Fatal CIGAR ops exit(1)
 * in the reference, so callers must not pass them here. */
int refh_get_mod_poss_on_ref(const uint32_t *cigar, int cigar_l, uint32_t qs, int strand, const uint32_t *mod_poss,
                             const uint8_t *mod_cat, int mod_l, const uint8_t *seqi, uint32_t aln_len,
                             uint32_t *out_pos, uint8_t *out_cat, int cap, int *n_out) {
    mod_t m;
    init_mod_t(&m, 16);
    uint32_t *mp = (uint32_t *)malloc(sizeof(uint32_t) * (mod_l + 2));
    uint8_t *mq = (uint8_t *)malloc(mod_l + 2);
    memcpy(mp, mod_poss, sizeof(uint32_t) * mod_l);
    memcpy(mq, mod_cat, mod_l);
    int rc = get_mod_poss_on_ref(&m, (uint32_t *)cigar, cigar_l, qs, strand, mp, mq, mod_l, (uint8_t *)seqi, aln_len,
                                 (char *)"q");
    *n_out = (int)m.calls.n;
    for (int i = 0; i < (int)m.calls.n && i < cap; i++) { out_pos[i] = m.calls.a[i]; out_cat[i] = m.quals.a[i]; }
    free(mp); free(mq);
    destroy_mod_t(&m, 0);
    return rc;
}

/* ---- -u pre-haplotagging through the reference's own loader (blockjoin.c:4446-4448) ---- */
typedef struct { storage_t *st; } refh_tags_t;

void *refh_pre_haplotag(const char *fn_vcf, const char *fn_bam) {
    storage_t *st = (storage_t *)calloc(1, sizeof(storage_t));
    init_storage_t(st);
    load_intervals_from_file((char *)fn_vcf, IS_VCF, st, 1, (char *)fn_bam, 0, 0);
    return st;
}
void *refh_tags_hash(void *st) { return ((storage_t *)st)->qname2haptag_raw; }
int refh_tags_count(void *stp) { return (int)kh_size(((storage_t *)stp)->qname2haptag_raw); }
int refh_tag_lookup(void *stp, const char *qname) {
    htstri_t *h = ((storage_t *)stp)->qname2haptag_raw;
    khint_t k = htstri_ht_get(h, (char *)qname);
    if (k == kh_end(h)) return -1;
    return kh_val(h, k);
}
void refh_tags_free(void *stp) { destroy_storage_t((storage_t *)stp, 1); }

/* known variants of one contig as insert_variant_from_vcf_line builds them */
int refh_load_variants(const char *fn_vcf, const char *chrom, uint32_t *pos, uint32_t *len, uint8_t *op,
                       uint8_t *haptag, uint8_t *bases, uint32_t *bases_off, int cap, int bases_cap) {
    vvar_t vars;
    init_vvar_t(&vars);
    ranges_t *r = load_intervals_from_file_one_ref((char *)fn_vcf, (char *)chrom, IS_VCF, &vars, -1, -1);
    destroy_ranges_t(r);
    int n = (int)vars.n, bo = 0;
    for (int i = 0; i < n && i < cap; i++) {
        pos[i] = vars.a[i].pos; len[i] = vars.a[i].len; op[i] = vars.a[i].op; haptag[i] = vars.a[i].haptag;
        bases_off[i] = (uint32_t)bo;
        for (size_t j = 0; j < vars.a[i].chars.n && bo < bases_cap; j++) bases[bo++] = vars.a[i].chars.a[j];
    }
    destroy_vvar_t(&vars, 0);
    return n;
}

/* ---- interval bookkeeping (a16) ---- */
typedef struct {
    int ref_n;
    storage_t *st;
} refh_intervals_t;

void *refh_load_intervals(const char *fn, int fmt /*0 gtf 1 vcf 2 tsv*/) {
    storage_t *st = (storage_t *)calloc(1, sizeof(storage_t));
    init_storage_t(st);
    load_intervals_from_file((char *)fn, (enum input_file_format)fmt, st, 0, 0, 0, 0);
    for (int i = 0; i < st->ref_n; i++) {
        store_raw_intervals(st->ranges[i]);
        merge_close_intervals(st->ranges[i], READBACK);
    }
    return st;
}
int refh_intervals_nref(void *stp) { return ((storage_t *)stp)->ref_n; }
const char *refh_intervals_refname(void *stp, int i) { return ((storage_t *)stp)->ref_names[i]; }
int refh_intervals_n(void *stp, int i) { return (int)((storage_t *)stp)->ranges[i]->starts.n; }
void refh_intervals_get(void *stp, int i, uint32_t *starts, uint32_t *ends, uint32_t *abs_se) {
    ranges_t *r = ((storage_t *)stp)->ranges[i];
    memcpy(starts, r->starts.a, sizeof(uint32_t) * r->starts.n);
    memcpy(ends, r->ends.a, sizeof(uint32_t) * r->starts.n);
    abs_se[0] = r->abs_start;
    abs_se[1] = r->abs_end;
}
/* apply decisions, then lift/flip/blocks (blockjoin.c:4677-4679); returns #phase blocks of ref i via getters */
void refh_intervals_decide(void *stp, int i, const int *decisions) {
    ranges_t *r = ((storage_t *)stp)->ranges[i];
    for (size_t j = 0; j < r->starts.n; j++) r->decisions.a[j] = decisions[j];
}
void refh_intervals_finish(void *stp) {
    storage_t *st = (storage_t *)stp;
    lift_decisions(st);
    make_decisions_flippings_onraw(st);
    generate_new_phase_blocks(st, 1);
}
int refh_intervals_nblocks(void *stp, int i) { return (int)((storage_t *)stp)->ranges[i]->phaseblocks.n; }
void refh_intervals_blocks(void *stp, int i, uint32_t *s, uint32_t *e) {
    ranges_t *r = ((storage_t *)stp)->ranges[i];
    for (size_t j = 0; j < r->phaseblocks.n; j++) { s[j] = r->phaseblocks.a[j].s; e[j] = r->phaseblocks.a[j].e; }
}
void refh_intervals_free(void *stp) { destroy_storage_t((storage_t *)stp, 1); }

double refh_fisher_two_sided(int a, int b, int c, int d) {
    double l, r, t;
    kt_fisher_exact(a, b, c, d, &l, &r, &t);
    return t;
}
float refh_evaluate_separation1(const uint8_t *ref, const uint8_t *query, int n, int *join_dir) {
    return evaluate_separation1((uint8_t *)ref, (uint8_t *)query, n, join_dir, NULL);
}

/* ---- one hand-built record through the reference's fill_read_meth_record_from_bam_line (blockjoin.c:794-908)
 * on top of the shim's bam_parse_basemod / bam_mods_at_next_pos: known-answer vectors for the MM/ML layer.
 * mm == NULL: no MM tag; ml_len < 0: no ML tag; mn < 0: no MN tag.  Returns the function's return value. */
int refh_fill_read_meth(uint32_t pos, int flag, const uint32_t *cigar, int n_cigar, const uint8_t *seq4, int l_qseq,
                        const char *mm, const uint8_t *ml, int ml_len, int mn, int lo, int hi, uint32_t *out_pos,
                        uint8_t *out_cat, int cap, int *n_out, int *has_implicit) {
    bam1_t *b = bam_init1();
    size_t mm_l = mm ? strlen(mm) : 0;
    size_t need = 4 + (size_t)n_cigar * 4 + ((size_t)l_qseq + 1) / 2 + (size_t)l_qseq + 3 + mm_l + 1 + 8 + (ml_len > 0 ? ml_len : 0) + 16;
    b->data = (uint8_t *)calloc(1, need);
    b->m_data = (uint32_t)need;
    uint8_t *p = b->data;
    memcpy(p, "q\0\0\0", 4); p += 4;
    b->core.l_qname = 4; b->core.l_extranul = 2;
    memcpy(p, cigar, (size_t)n_cigar * 4); p += (size_t)n_cigar * 4;
    memcpy(p, seq4, ((size_t)l_qseq + 1) / 2); p += ((size_t)l_qseq + 1) / 2;
    memset(p, 0xff, (size_t)l_qseq); p += l_qseq;
    if (mm) { *p++ = 'M'; *p++ = 'M'; *p++ = 'Z'; memcpy(p, mm, mm_l + 1); p += mm_l + 1; }
    if (ml_len >= 0) {
        *p++ = 'M'; *p++ = 'L'; *p++ = 'B'; *p++ = 'C';
        uint32_t n = (uint32_t)ml_len;
        memcpy(p, &n, 4); p += 4;
        if (ml_len) memcpy(p, ml, (size_t)ml_len);
        p += ml_len;
    }
    if (mn >= 0) { *p++ = 'M'; *p++ = 'N'; *p++ = 'I'; uint32_t v = (uint32_t)mn; memcpy(p, &v, 4); p += 4; }
    b->l_data = (int)(p - b->data);
    b->core.pos = (int32_t)pos; b->core.flag = (uint16_t)flag; b->core.n_cigar = (uint32_t)n_cigar; b->core.l_qseq = l_qseq;
    b->core.qual = 60; b->core.tid = 0; b->core.mtid = -1; b->core.mpos = -1;
    read_t h;
    memset(&h, 0, sizeof(h));
    init_mod_t(&h.meth, 16);
    hts_base_mod_state *ms = hts_base_mod_state_alloc();
    vu32_t poss; vu8_t quals;
    kv_init(poss); kv_init(quals);
    global_data_has_implicit = 0;
    int stat = fill_read_meth_record_from_bam_line(&h, b, ms, &poss, &quals, (uint8_t)lo, (uint8_t)hi);
    *n_out = (int)h.meth.calls.n;
    for (int i = 0; i < (int)h.meth.calls.n && i < cap; i++) { out_pos[i] = h.meth.calls.a[i]; out_cat[i] = h.meth.quals.a[i]; }
    if (has_implicit) *has_implicit = global_data_has_implicit;
    kv_destroy(poss); kv_destroy(quals);
    destroy_mod_t(&h.meth, 0);
    hts_base_mod_state_free(ms);
    bam_destroy1(b);
    return stat;
}
