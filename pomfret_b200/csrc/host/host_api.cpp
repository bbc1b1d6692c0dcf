// C entry points of the host front end used by the Python tests and bench.py (ctypes).
#include <cstring>
#include "loader.h"

using namespace pomfret;

extern "C" {

void *pomfret_host_bam_open(const char *path) {
    BamReader *r = new BamReader();
    if (!r->open(path)) { delete r; return nullptr; }
    return r;
}
void pomfret_host_bam_close(void *h) { delete (BamReader *)h; }

void *pomfret_host_window_load(void *bam, const char *chrom, uint32_t ref_start, uint32_t ref_end,
                               int readlen_threshold, int min_mapq, int *rc_out) {
    WindowReads *w = new WindowReads();
    int rc = load_window(*(BamReader *)bam, chrom, ref_start, ref_end, readlen_threshold, min_mapq, nullptr, w);
    if (rc_out) *rc_out = rc;
    if (rc != 0) { delete w; return nullptr; }
    return w;
}
int pomfret_host_window_n(void *w) { return (int)((WindowReads *)w)->descs.size(); }
const pomfret_gpu_read_desc *pomfret_host_window_descs(void *w) { return ((WindowReads *)w)->descs.data(); }
const char *pomfret_host_window_qname(void *w, int i) { return ((WindowReads *)w)->qname((size_t)i); }
uint64_t pomfret_host_window_bases(void *w) { return ((WindowReads *)w)->n_bases; }
// the buffer that holds the window's record copies (what the descriptors point into)
const uint8_t *pomfret_host_window_arena(void *w, uint64_t *n_bytes) {
    WindowReads *r = (WindowReads *)w;
    if (n_bytes) *n_bytes = r->arena.size();
    return r->arena.data();
}
void pomfret_host_window_free(void *w) { delete (WindowReads *)w; }

}  // extern "C"

// ---- interval bookkeeping hooks (tests compare them with the compiled reference) ----
#include "intervals.h"
#include "methphase.h"
extern "C" {

void *pomfret_host_intervals_load(const char *fn, int fmt) {
    Storage *st = new Storage();
    std::string fatal;
    if (!load_intervals(fn, (IntervalFormat)fmt, st, nullptr, &fatal) || !fatal.empty()) { delete st; return nullptr; }
    for (Ranges &r : st->ranges) { store_raw_intervals(&r); merge_close_intervals(&r, kReadback); }
    return st;
}
int pomfret_host_intervals_nref(void *h) { return (int)((Storage *)h)->ref_names.size(); }
const char *pomfret_host_intervals_refname(void *h, int i) { return ((Storage *)h)->ref_names[(size_t)i].c_str(); }
int pomfret_host_intervals_n(void *h, int i) { return (int)((Storage *)h)->ranges[(size_t)i].n; }
void pomfret_host_intervals_get(void *h, int i, uint32_t *starts, uint32_t *ends, uint32_t *abs_se) {
    const Ranges &r = ((Storage *)h)->ranges[(size_t)i];
    for (size_t j = 0; j < r.n; j++) { starts[j] = r.starts[j]; ends[j] = r.ends[j]; }
    abs_se[0] = r.abs_start; abs_se[1] = r.abs_end;
}
void pomfret_host_intervals_decide(void *h, int i, const int *decisions) {
    Ranges &r = ((Storage *)h)->ranges[(size_t)i];
    for (size_t j = 0; j < r.n; j++) r.decisions[j] = decisions[j];
}
void pomfret_host_intervals_finish(void *h) {
    Storage *st = (Storage *)h;
    lift_decisions(st);
    make_flips_onraw(st);
    generate_new_phase_blocks(st);
}
int pomfret_host_intervals_nblocks(void *h, int i) { return (int)((Storage *)h)->ranges[(size_t)i].phaseblocks.size(); }
void pomfret_host_intervals_blocks(void *h, int i, uint32_t *s, uint32_t *e) {
    const Ranges &r = ((Storage *)h)->ranges[(size_t)i];
    for (size_t j = 0; j < r.phaseblocks.size(); j++) { s[j] = r.phaseblocks[j].s; e[j] = r.phaseblocks[j].e; }
}
void pomfret_host_intervals_free(void *h) { delete (Storage *)h; }

// known variants of one contig as the -u path sees them
int pomfret_host_load_variants(const char *fn_vcf, const char *chrom, pomfret_gpu_variant *vars, int cap, uint8_t *bases, int bases_cap,
                               int *n_bases) {
    Storage st;
    std::string fatal;
    int n = 0, nb = 0;
    load_intervals(fn_vcf, IntervalFormat::VCF, &st,
                   [&](const std::string &c, KnownVariants &kv, bool) {
                       if (c != chrom) return;
                       for (size_t i = 0; i < kv.vars.size(); i++) {
                           if (n < cap) {
                               vars[n] = kv.vars[i];
                               vars[n].bases_off = (uint32_t)nb;
                               for (uint32_t j = 0; j < kv.vars[i].len && nb < bases_cap; j++) bases[nb++] = kv.bases[kv.vars[i].bases_off + j];
                           }
                           n++;
                       }
                   }, &fatal);
    if (n_bases) *n_bases = nb;
    return n;
}

// every primary record of a contig, packed for the haplotagger (tests drive pomfret_gpu_haptag with these)
void *pomfret_host_contig_load(void *bam, const char *chrom) {
    BamReader &b = *(BamReader *)bam;
    WindowReads *w = new WindowReads();
    hts_itr_t *itr = sam_itr_querys(b.idx, b.hdr, chrom);
    if (!itr) return w;
    std::vector<size_t> offs, lens;
    std::vector<bam1_core_t> cores;
    while (sam_itr_next(b.fp, itr, b.rec) >= 0) {
        const int flag = b.rec->core.flag;
        if ((flag & 4) || (flag & 256) || (flag & 2048)) continue;
        size_t off = (w->arena.size() + 15) & ~(size_t)15;
        w->arena.resize(off + (size_t)b.rec->l_data + 16);
        memcpy(w->arena.data() + off, b.rec->data, (size_t)b.rec->l_data);
        offs.push_back(off); lens.push_back((size_t)b.rec->l_data); cores.push_back(b.rec->core);
        w->qname_off.push_back((uint32_t)w->qnames.size());
        w->qnames.append(bam_get_qname(b.rec));
        w->qnames.push_back('\0');
        w->n_bases += (uint64_t)b.rec->core.l_qseq;
    }
    hts_itr_destroy(itr);
    w->descs.resize(offs.size());
    bam1_t tmp;
    memset(&tmp, 0, sizeof(tmp));
    for (size_t i = 0; i < offs.size(); i++) {
        tmp.core = cores[i];
        tmp.data = w->arena.data() + offs[i];
        tmp.l_data = (int)lens[i];
        describe_record(&tmp, kHaptagUnphased, &w->descs[i]);
    }
    return w;
}

}  // extern "C"

// ---- compressed ingest: the block / stream tables of region queries, for the tests and bench.py ----
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include "ingest.h"
extern "C" {

struct HostIngest {
    IngestPlan plan;
    std::vector<uint8_t> comp;
};

// regions: n x (beg0, end0) on `chrom`, one run each; the compressed blocks are read into the handle's own buffer
void *pomfret_host_ingest_plan(void *bam, const char *chrom, const int64_t *regions, int n) {
    BamReader &b = *(BamReader *)bam;
    const int tid = sam_hdr_name2tid(b.hdr, chrom);
    if (tid < 0) return nullptr;
    int fd = open(b.fn.c_str(), O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) return nullptr;
    HostIngest *h = new HostIngest();
    bool ok = true;
    for (int i = 0; i < n && ok; i++) ok = ingest_plan_region(b.idx, tid, regions[2 * i], regions[2 * i + 1], (uint32_t)i, (uint64_t)st.st_size, &h->plan);
    h->comp.assign(h->plan.comp_bytes + 64, 0);
    std::string err;
    if (ok) ok = ingest_read(fd, &h->plan, h->comp.data(), &err);
    close(fd);
    if (!ok) { delete h; return nullptr; }
    return h;
}
const uint8_t *pomfret_host_ingest_comp(void *h, uint64_t *n) { HostIngest *p = (HostIngest *)h; *n = p->plan.comp_bytes; return p->comp.data(); }
const pomfret_gpu_bgzf_block *pomfret_host_ingest_blocks(void *h, uint32_t *n) { HostIngest *p = (HostIngest *)h; *n = (uint32_t)p->plan.blocks.size(); return p->plan.blocks.data(); }
const pomfret_gpu_bgzf_stream *pomfret_host_ingest_streams(void *h, uint32_t *n) { HostIngest *p = (HostIngest *)h; *n = (uint32_t)p->plan.streams.size(); return p->plan.streams.data(); }
uint32_t pomfret_host_ingest_stream_run(void *h, uint32_t s) { return ((HostIngest *)h)->plan.stream_run[s]; }
void pomfret_host_ingest_free(void *h) { delete (HostIngest *)h; }

}  // extern "C"
