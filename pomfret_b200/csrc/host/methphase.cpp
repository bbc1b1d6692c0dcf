// Drivers, see methphase.h.
#include "methphase.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <dirent.h>
#include <future>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_set>
#include "gpu_api.h"
#include "intervals.h"
#include "loader.h"
#include "ingest.h"
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace pomfret {

namespace {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

[[noreturn]] void die_gpu(const GpuApi &api, int rc, const char *where) {
    if (rc == POMFRET_GPU_ERR_FATAL_CIGAR)
        fprintf(stderr, "[E::%s] fatal: unknown cigar operation. Bug?\n", "get_mod_poss_on_ref");  // blockjoin.c:777
    else if (rc == POMFRET_GPU_ERR_MISSING_MD)
        fprintf(stderr, "pomfret: parse_variants_for_one_read: Assertion `tagd' failed.\n");       // blockjoin.c:1596
    else if (rc == POMFRET_GPU_ERR_BAD_MD)
        fprintf(stderr, "[E::%s] invalid MD\n", "parse_variants_for_one_read");                    // blockjoin.c:1622
    else
        fprintf(stderr, "[E::%s] %s: %s\n", "pomfret_gpu", where, api.strerror ? api.strerror(rc) : "?");
    exit(1);
}

struct Engine {
    GpuApi &api = gpu_api();
    pomfret_gpu_ctx *ctx = nullptr;
    int n_dev = 0;
    bool start(int want_gpus, int n_workers) {
        const double t0 = now_s();
        std::string err;
        if (!api.load(&err)) { fprintf(stderr, "[E::%s] %s\n", "pomfret", err.c_str()); return false; }
        int n = api.device_count();
        if (n <= 0) { fprintf(stderr, "[E::%s] no CUDA device visible; this build has no CPU path\n", "pomfret"); return false; }
        n_dev = want_gpus > 0 && want_gpus < n ? want_gpus : n;
        std::vector<int> devs;
        for (int i = 0; i < n_dev; i++) devs.push_back(i);
        int rc = api.init(&ctx, devs.data(), n_dev, n_workers);
        if (rc != 0) { fprintf(stderr, "[E::%s] pomfret_gpu_init: %s\n", "pomfret", api.strerror(rc)); return false; }
        fprintf(stderr, "[T::%s] engine ready after %.2fs (%d devices)\n", "pomfret", now_s() - t0, n_dev);
        return true;
    }
    // CUDA start-up (driver, contexts: about half a second on a B200 box) runs beside the host work that does not
    // need the device: interval loading, BAM/index opening and the inflate of every worker's first chunk
    // compressed ingest (BGZF inflate + record slicing on the device) unless POMFRET_HOST_INFLATE is set
    bool gpu_ingest = getenv("POMFRET_HOST_INFLATE") == nullptr;
    std::shared_future<bool> ready;
    void start_async(int want_gpus, int n_workers) {
        ready = std::async(std::launch::async, [this, want_gpus, n_workers] { return start(want_gpus, n_workers); }).share();
    }
    bool wait() { return ready.valid() ? ready.get() : ctx != nullptr; }
    bool is_ready() { return !ready.valid() || ready.wait_for(std::chrono::seconds(0)) == std::future_status::ready; }
    // devices the engine will report, without waiting for it: --gpus, else what CUDA_VISIBLE_DEVICES / the driver list
    int expected_devices(int want_gpus) {
        if (ready.valid() && ready.wait_for(std::chrono::seconds(0)) == std::future_status::ready) return ready.get() ? n_dev : 1;
        int n = 0;
        if (const char *v = getenv("CUDA_VISIBLE_DEVICES")) {
            if (*v) { n = 1; for (const char *p = v; *p; p++) if (*p == ',') n++; }
        } else if (DIR *d = opendir("/proc/driver/nvidia/gpus")) {
            while (struct dirent *e = readdir(d)) if (e->d_name[0] != '.') n++;
            closedir(d);
        }
        if (n <= 0) n = 1;
        return want_gpus > 0 && want_gpus < n ? want_gpus : n;
    }
    ~Engine() { if (ready.valid()) ready.wait(); if (ctx) api.destroy(ctx); }
};

struct WindowJob {
    int i_ref;
    size_t i_win;
    uint32_t start, end;
    // the region query of load_reads_given_interval (blockjoin.c:1053-1061) as the iterator sees it: [beg0, end0)
    int64_t beg0() const { int64_t b = (int64_t)start - kReadback; if (b < 0) b = 0; b -= 1; return b < 0 ? 0 : b; }
    int64_t end0() const { return (int64_t)end + kReadback; }
};

struct WindowOut {
    int decision = -1;
    std::vector<std::pair<std::string, int>> tags;  // kept reads in BAM order, only when decision >= 0
};

// Record storage of one worker: slabs of host memory the BAM reader inflates records straight into.  add_reads()
// copies the fields the device needs (a third of a record: base qualities stay behind) into the engine's pinned
// arena.  (Registering the slabs instead — pomfret_gpu_host_register, no copy — pins three times the bytes, and
// pinning costs more than the copy until a run is long enough to reuse the slabs many times: POMFRET_PIN_RECORDS=1.)
struct RecordArena {
    static constexpr size_t kSlab = (size_t)64 << 20, kMinFree = (size_t)24 << 20;
    bool pin = getenv("POMFRET_PIN_RECORDS") != nullptr;
    struct Slab { uint8_t *p; size_t cap, len; };
    std::vector<Slab> slabs;
    size_t cur = 0;
    const GpuApi *api = nullptr;
    pomfret_gpu_ctx *ctx = nullptr;
    void rewind() { for (Slab &s : slabs) s.len = 0; cur = 0; }
    // a place for one record of at most kMinFree bytes
    uint8_t *room(size_t *cap) {
        if (pin && !api) pin = false;  // (the engine is still starting: these slabs stay unpinned)
        while (cur < slabs.size() && slabs[cur].cap - slabs[cur].len < kMinFree) cur++;
        if (cur == slabs.size()) {
            void *p = nullptr;
            if (posix_memalign(&p, 4096, kSlab) != 0) return nullptr;
            if (pin) {
                int rc = api->host_register(ctx, p, kSlab);
                if (rc != 0) { fprintf(stderr, "[W::%s] host_register: %s (records will be copied instead)\n", "pomfret", api->strerror(rc)); pin = false; }
            }
            slabs.push_back({(uint8_t *)p, kSlab, 0});
        }
        *cap = slabs[cur].cap - slabs[cur].len;
        return slabs[cur].p + slabs[cur].len;
    }
    void commit(size_t n) { slabs[cur].len += (n + 63) & ~(size_t)63; }
    void release() {
        for (Slab &s : slabs) { if (pin) api->host_unregister(ctx, s.p); free(s.p); }
        slabs.clear();
    }
};

// One worker: its own BAM handle, record arena and batch on its own device/stream.
struct Worker {
    Engine *eng = nullptr;
    int id = 0, device = 0;
    BamReader bam;
    pomfret_gpu_batch *batch = nullptr;
    RecordArena arena;
    bam1_t view;   // record header over arena memory (POMFRET_BAM_EXTERNAL_DATA)
    RunStats stats;
    bool open(const std::string &fn_bam) {
        if (!bam.open(fn_bam)) { fprintf(stderr, "[E::%s] failed to open input file: %s\n", "load_reads_given_interval", fn_bam.c_str()); return false; }
        memset(&view, 0, sizeof(view));
        view.id = POMFRET_BAM_EXTERNAL_DATA;
        return true;
    }
    // the device side of the worker: waits for the engine's start-up if that is still under way
    void need_batch() {
        if (batch) return;
        if (!eng->wait()) exit(1);
        arena.api = &eng->api; arena.ctx = eng->ctx;
        int rc = eng->api.batch_begin(eng->ctx, id, device % std::max(1, eng->n_dev), &batch);
        if (rc != 0) { fprintf(stderr, "[E::%s] batch_begin: %s\n", "pomfret", eng->api.strerror(rc)); exit(1); }
    }
    void close() { if (batch) eng->api.batch_end(batch); batch = nullptr; arena.release(); if (fd >= 0) ::close(fd); fd = -1; }

    // next record of the iterator, inflated in place into the arena; false at the end
    bool next_record(hts_itr_t *itr) {
        size_t cap = 0;
        view.data = arena.room(&cap);
        if (!view.data) { fprintf(stderr, "[E::%s] out of host memory\n", "pomfret"); exit(1); }
        view.m_data = (uint32_t)std::min<size_t>(cap, 0xffffffffu);
        view.l_data = 0;
        const int rc = sam_itr_next(bam.fp, itr, &view);
        if (rc < -1) { fprintf(stderr, "[E::%s] error while reading %s (truncated file, or a record larger than %zu MB)\n", "pomfret", bam.fn.c_str(), RecordArena::kMinFree >> 20); exit(1); }
        return rc >= 0;
    }

    // one alignment record of a chunk, however it was loaded: its descriptor (host pointers into the arena, or device
    // addresses inside the inflated streams) and its name
    struct Rec { pomfret_gpu_read_desc desc; const char *qname; int64_t first_slot; };
    std::vector<Rec> recs;
    std::vector<std::vector<uint32_t>> win_recs;
    std::vector<char> names;   // device path: the names of the chunk's records
    int fd = -1;               // device path: the BAM file for pread()
    uint64_t file_size = 0;
    IngestPlan plan;

    // runs of windows whose region queries overlap: [r0, r1) index ranges into jobs
    static std::vector<std::pair<size_t, size_t>> runs_of(const std::vector<WindowJob> &jobs) {
        std::vector<std::pair<size_t, size_t>> runs;
        for (size_t r0 = 0; r0 < jobs.size();) {
            size_t r1 = r0 + 1;
            int64_t run_end = jobs[r0].end0();
            while (r1 < jobs.size() && jobs[r1].beg0() < run_end) { run_end = std::max(run_end, jobs[r1].end0()); r1++; }
            runs.emplace_back(r0, r1);
            r0 = r1;
        }
        return runs;
    }
    static int64_t run_end0(const std::vector<WindowJob> &jobs, const std::pair<size_t, size_t> &run) {
        int64_t e = 0;
        for (size_t w = run.first; w < run.second; w++) e = std::max(e, jobs[w].end0());
        return e;
    }

    // host loader: the shim inflates every run once, records land in the arena
    void load_chunk_host(const std::string &chrom, const std::vector<WindowJob> &jobs, const pomfret_gpu_config &cfg, const RawTagMap *raw_tags) {
        arena.rewind();
        const int tid = sam_hdr_name2tid(bam.hdr, chrom.c_str());
        for (const auto &run : runs_of(jobs)) {
            hts_itr_t *itr = tid >= 0 ? sam_itr_queryi(bam.idx, tid, jobs[run.first].beg0(), run_end0(jobs, run)) : nullptr;
            if (!itr) { fprintf(stderr, "[E::%s] region query failed for %s:%u-%u\n", "load_reads_given_interval", chrom.c_str(), jobs[run.first].start, jobs[run.first].end); exit(1); }
            while (next_record(itr)) {
                const bam1_t *b = &view;
                const int flag = b->core.flag;
                const uint32_t len = (uint32_t)b->core.l_qseq;
                if ((flag & 4) || (flag & 256) || (flag & 2048)) continue;      // blockjoin.c:1081-1084
                if (b->core.qual < (uint32_t)cfg.min_mapq) continue;
                if (len < 2 || len < (uint32_t)cfg.readlen_threshold) continue;
                float de = -1;
                if (uint8_t *t = bam_aux_get(b, "de")) de = (float)bam_aux2f(t);
                if (de > kMinAlnDe) continue;
                const int64_t pos = b->core.pos, endpos = bam_endpos(b);
                bool used = false;
                for (size_t w = run.first; w < run.second; w++) {
                    if (!(pos < jobs[w].end0() && endpos > jobs[w].beg0())) continue;  // what this window's own query returns
                    if (!used) {
                        int hp;
                        if (raw_tags) {
                            auto it = raw_tags->find(bam_get_qname(b));
                            hp = it != raw_tags->end() ? it->second : kHaptagUnphased;
                        } else hp = hp_from_record(b);
                        Rec R;
                        describe_record(b, hp, &R.desc);
                        R.desc.md = nullptr; R.desc.md_len = 0;  // the window engine does not read MD
                        R.qname = bam_get_qname(b);
                        R.first_slot = -1;
                        recs.push_back(R);
                        arena.commit((size_t)b->l_data);
                        used = true;
                    }
                    win_recs[w].push_back((uint32_t)recs.size() - 1);
                }
            }
            hts_itr_destroy(itr);
        }
    }

    // device loader (compressed ingest): the BGZF blocks of every run's index chunks go to the device as they lie in
    // the file; inflate, record walk, filters and tag lookup happen there; the host gets one header per record back
    // which blocks the chunk's region queries need (the index chunks of every run)
    void plan_chunk(const std::string &chrom, const std::vector<WindowJob> &jobs, IngestPlan *pl) {
        const int tid = sam_hdr_name2tid(bam.hdr, chrom.c_str());
        if (fd < 0) {
            fd = ::open(bam.fn.c_str(), O_RDONLY);
            struct stat st;
            if (fd < 0 || fstat(fd, &st) != 0) { fprintf(stderr, "[E::%s] failed to open input file: %s\n", "load_reads_given_interval", bam.fn.c_str()); exit(1); }
            file_size = (uint64_t)st.st_size;
        }
        pl->clear();
        const auto runs = runs_of(jobs);
        for (size_t r = 0; r < runs.size(); r++)
            if (tid < 0 || !ingest_plan_region(bam.idx, tid, jobs[runs[r].first].beg0(), run_end0(jobs, runs[r]), (uint32_t)r, file_size, pl)) {
                fprintf(stderr, "[E::%s] region query failed for %s:%u-%u\n", "load_reads_given_interval", chrom.c_str(), jobs[runs[r].first].start, jobs[runs[r].first].end);
                exit(1);
            }
    }
    void read_chunk(IngestPlan *pl, uint8_t *comp) {
        std::string err;
        if (!ingest_read(fd, pl, comp, &err)) { fprintf(stderr, "[E::%s] %s: %s\n", "pomfret", bam.fn.c_str(), err.c_str()); exit(1); }
    }
    // A chunk whose blocks were read ahead of time (while the CUDA start-up was still under way)
    struct ReadAhead { size_t chunk; IngestPlan plan; std::vector<uint8_t> comp; };
    const ReadAhead *ahead = nullptr;
    std::vector<uint8_t> comp_buf;  // the blocks of the chunk at hand

    void load_chunk_device(const std::string &chrom, const std::vector<WindowJob> &jobs, const pomfret_gpu_config &cfg, const RawTagMap *raw_tags) {
        const GpuApi &api = eng->api;
        double tp = now_s();
        auto lap = [&](int i) { const double t = now_s(); stats.t_phase[i] += t - tp; tp = t; };
        if (ahead) plan = ahead->plan; else plan_chunk(chrom, jobs, &plan);
        const auto runs = runs_of(jobs);
        lap(0);
        need_batch();
        lap(1);
        int rc;
        if ((rc = api.batch_reset(batch))) die_gpu(api, rc, "batch_reset");
        lap(2);
        // (plain memory: the blocks cross the bus once, pinning a buffer for them costs more than the driver's staging copy)
        const void *comp = nullptr;
        if (ahead) comp = ahead->comp.data();
        else {
            if (comp_buf.size() < plan.comp_bytes + 64) comp_buf.resize(plan.comp_bytes + plan.comp_bytes / 4 + 64);
            read_chunk(&plan, comp_buf.data());
            comp = comp_buf.data();
        }
        lap(3);
        pomfret_gpu_ingest_filter flt;
        memset(&flt, 0, sizeof(flt));
        flt.min_mapq = (uint32_t)std::max(0, cfg.min_mapq);
        flt.min_len = (uint32_t)std::max(0, cfg.readlen_threshold);
        flt.min_len_floor = 2; flt.check_de = 1; flt.max_de = kMinAlnDe;
        uint32_t n_rec = 0;
        if ((rc = api.batch_ingest_bgzf(batch, comp, plan.comp_bytes, plan.blocks.data(), (uint32_t)plan.blocks.size(), plan.streams.data(),
                                        (uint32_t)plan.streams.size(), &flt, &n_rec)))
            die_gpu(api, rc, "ingest_bgzf");
        std::vector<pomfret_gpu_sliced_record> sl(n_rec ? n_rec : 1);
        lap(4);
        if ((rc = api.batch_ingest_records(batch, sl.data(), n_rec))) die_gpu(api, rc, "ingest_records");
        lap(5);
        stats.n_ingest_bytes += plan.comp_bytes;
        names.clear();
        names.reserve((size_t)n_rec * 40);
        std::vector<size_t> name_off;
        for (uint32_t i = 0; i < n_rec; i++) {
            const pomfret_gpu_sliced_record &S = sl[i];
            if (S.bad) { fprintf(stderr, "[E::%s] malformed alignment record in %s\n", "pomfret", bam.fn.c_str()); exit(1); }
            if (!S.keep) continue;
            const auto &run = runs[plan.stream_run[S.stream]];
            bool used = false;
            for (size_t w = run.first; w < run.second; w++) {
                if (!((int64_t)S.pos < jobs[w].end0() && (int64_t)S.end_pos > jobs[w].beg0())) continue;
                if (!used) {
                    // the whole name (records with names longer than the inline prefix are fetched one by one)
                    name_off.push_back(names.size());
                    if (S.l_qname <= sizeof(S.qname)) names.insert(names.end(), S.qname, S.qname + strlen(S.qname) + 1);
                    else {
                        char buf[256];
                        if ((rc = api.batch_ingest_qname(batch, i, buf, sizeof(buf)))) die_gpu(api, rc, "ingest_qname");
                        names.insert(names.end(), buf, buf + strlen(buf) + 1);
                    }
                    const char *qn = names.data() + name_off.back();
                    int hp = S.hp;
                    if (raw_tags) {
                        auto it = raw_tags->find(qn);
                        hp = it != raw_tags->end() ? it->second : kHaptagUnphased;
                    } else if (S.hp_irregular)
                        fprintf(stderr, "[W::%s] irregular HP tag? qn=%s qs=%d\n", "get_hp_from_aln", qn, (int)S.pos);
                    Rec R;
                    pomfret_gpu_read_desc &d = R.desc;
                    memset(&d, 0, sizeof(d));
                    d.pos = S.pos; d.l_qseq = S.l_qseq; d.n_cigar = S.n_cigar; d.flag = S.flag; d.mapq = S.mapq;
                    d.tags_malformed = S.tags_malformed; d.hp = hp; d.mn = S.mn;
                    d.cigar = (const uint32_t *)(uintptr_t)S.cigar; d.seq = (const uint8_t *)(uintptr_t)S.seq;
                    d.mm = S.has_mm ? (const char *)(uintptr_t)S.mm : nullptr; d.mm_len = S.mm_len;
                    d.ml = (const uint8_t *)(uintptr_t)S.ml; d.ml_len = S.ml_len;
                    d.reserved = S.end_pos;
                    R.qname = nullptr;  // set below: `names` may still move
                    R.first_slot = -1;
                    recs.push_back(R);
                    used = true;
                }
                win_recs[w].push_back((uint32_t)recs.size() - 1);
            }
        }
        for (size_t i = 0; i < recs.size(); i++) recs[i].qname = names.data() + name_off[i];
        lap(6);
    }

    // haplotag_region_given_bam for a chunk of windows of one contig.  Windows whose region queries overlap
    // form a run that is read from the BAM once; a record that lies in several windows is staged and decoded
    // once (the reference re-opens the file and re-decodes per window, blockjoin.c:1056).
    void run_chunk(const std::string &chrom, const std::vector<WindowJob> &jobs, const pomfret_gpu_config &cfg,
                   const RawTagMap *raw_tags, std::vector<WindowOut> *outs) {
        const GpuApi &api = eng->api;
        double t0 = now_s();
        int rc;
        recs.clear();
        win_recs.assign(jobs.size(), {});
        const bool on_device = eng->gpu_ingest;
        if (on_device) load_chunk_device(chrom, jobs, cfg, raw_tags);
        else load_chunk_host(chrom, jobs, cfg, raw_tags);
        // slots in window order; the second and later uses of a record share the first one's payload and calls
        std::vector<pomfret_gpu_read_desc> descs;
        std::vector<int64_t> same_as;
        std::vector<uint32_t> w_start, w_end, w_first, w_n;
        bool any_shared = false;
        for (size_t w = 0; w < jobs.size(); w++) {
            w_start.push_back(jobs[w].start); w_end.push_back(jobs[w].end);
            w_first.push_back((uint32_t)descs.size()); w_n.push_back((uint32_t)win_recs[w].size());
            for (uint32_t ri : win_recs[w]) {
                Rec &R = recs[ri];
                if (R.first_slot < 0) {
                    R.first_slot = (int64_t)descs.size();
                    same_as.push_back(-1);
                    stats.n_bases += R.desc.l_qseq;
                    descs.push_back(R.desc);
                } else {
                    pomfret_gpu_read_desc d;
                    memset(&d, 0, sizeof(d));
                    d.pos = R.desc.pos; d.l_qseq = R.desc.l_qseq; d.n_cigar = R.desc.n_cigar; d.hp = R.desc.hp;
                    d.mn = -1; d.ml_len = -1;
                    same_as.push_back(R.first_slot);
                    any_shared = true;
                    stats.n_shared++;
                    descs.push_back(d);
                }
            }
            stats.n_reads += win_recs[w].size();
        }
        if (!on_device) {
            need_batch();
            if ((rc = api.batch_reset(batch))) die_gpu(api, rc, "batch_reset");
        }
        if (!descs.empty()) {
            rc = on_device ? api.batch_add_reads_device(batch, descs.data(), (uint32_t)descs.size(), any_shared ? same_as.data() : nullptr)
                           : api.batch_add_reads_shared(batch, descs.data(), (uint32_t)descs.size(), any_shared ? same_as.data() : nullptr);
            if (rc) die_gpu(api, rc, "batch_add_reads");
        }
        if ((rc = api.batch_add_windows(batch, w_start.data(), w_end.data(), w_first.data(), w_n.data(), (uint32_t)jobs.size()))) die_gpu(api, rc, "batch_add_window");
        stats.n_windows += jobs.size();
        double t1 = now_s();
        stats.t_load += t1 - t0;
        if ((rc = api.batch_submit(batch))) die_gpu(api, rc, "batch_submit");
        if ((rc = api.decode(batch, (uint8_t)cfg.lo, (uint8_t)cfg.hi))) die_gpu(api, rc, "decode");
        stats.t_phase[7] += now_s() - t1;
        if ((rc = api.pileup(batch, &cfg))) die_gpu(api, rc, "pileup");
        stats.t_phase[8] += now_s() - t1;
        if ((rc = api.join(batch, &cfg))) die_gpu(api, rc, "join");
        const size_t n_slots = descs.size();
        std::vector<pomfret_gpu_window_result> res(jobs.size() ? jobs.size() : 1);
        std::vector<uint8_t> tags(n_slots ? n_slots : 1);
        std::vector<int32_t> ids(n_slots ? n_slots : 1);
        if ((rc = api.batch_collect(batch, res.data(), tags.data(), ids.data()))) die_gpu(api, rc, "batch_collect");
        stats.t_gpu += now_s() - t1;
        outs->assign(jobs.size(), WindowOut());
        for (size_t w = 0; w < jobs.size(); w++) {
            const size_t n = win_recs[w].size();
            auto qname = [&](size_t i) { return recs[win_recs[w][i]].qname; };
            // duplicated read names among the loaded records are fatal (blockjoin.c:1143-1155)
            std::unordered_set<std::string> seen;
            for (size_t i = 0; i < n; i++) {
                if (ids[w_first[w] + i] < 0) continue;
                if (!seen.insert(qname(i)).second) {
                    fprintf(stderr, "[E::%s] duplicated read name seen from reading bam: %s\n", "load_reads_given_interval", qname(i));
                    exit(1);
                }
            }
            WindowOut &o = (*outs)[w];
            o.decision = res[w].decision;
            fprintf(stderr, "[dbg::%s] loaded %d reads (interval: %s:%u-%u), left has #ref=%d (strict: %d), right has=%d (strict: %d); "
                            "sites n=%d; fwd join %d, bwd join %d\n", "haplotag_region_given_bam", res[w].n_reads, chrom.c_str(), jobs[w].start,
                    jobs[w].end, res[w].n_left, res[w].n_left_strict, res[w].n_right, res[w].n_right_strict, res[w].n_sites_fwd,
                    res[w].join_fwd, res[w].join_bwd);
            if (o.decision >= 0)
                for (size_t i = 0; i < n; i++)
                    if (ids[w_first[w] + i] >= 0) o.tags.emplace_back(qname(i), (int)tags[w_first[w] + i]);
        }
    }

    // recover_variant_phase_in_one_interval (blockjoin.c:2475-2600) on the device: the records of the interval are
    // inflated and sliced there, the host looks their names up (which reads did methylation phasing tag, and how),
    // haptag_kernel parses every record's own variants, variant_vote_kernel counts the votes per known position.
    void recover_interval_device(const PhaseState &ps, const std::string &chrom, uint32_t start, uint32_t end, const std::vector<uint32_t> &poss,
                                 std::unordered_map<uint32_t, uint32_t> *pos2hap) {
        const GpuApi &api = eng->api;
        if (poss.empty()) return;  // (the reference reads the interval anyway; nothing comes of it)
        need_batch();
        const int tid = sam_hdr_name2tid(bam.hdr, chrom.c_str());
        if (tid < 0) return;
        if (fd < 0) {
            fd = ::open(bam.fn.c_str(), O_RDONLY);
            struct stat st;
            if (fd < 0 || fstat(fd, &st) != 0) { fprintf(stderr, "[E::%s] failed to open input file: %s\n", "recover_variant_phase_in_one_interval", bam.fn.c_str()); exit(1); }
            file_size = (uint64_t)st.st_size;
        }
        plan.clear();
        // "%s:%d-%d" of the reference: 1-based inclusive start
        if (!ingest_plan_region(bam.idx, tid, start > 0 ? (int64_t)start - 1 : 0, (int64_t)end, 0, file_size, &plan)) return;
        int rc;
        if ((rc = api.batch_reset(batch))) die_gpu(api, rc, "batch_reset");
        if (plan.ranges.empty()) return;
        if (comp_buf.size() < plan.comp_bytes + 64) comp_buf.resize(plan.comp_bytes + plan.comp_bytes / 4 + 64);
        read_chunk(&plan, comp_buf.data());
        pomfret_gpu_ingest_filter flt;
        memset(&flt, 0, sizeof(flt));
        flt.keep_all_flags = 1;  // blockjoin.c:2512-2518: every record the query returns is looked up by name
        uint32_t n_rec = 0;
        if ((rc = api.batch_ingest_bgzf(batch, comp_buf.data(), plan.comp_bytes, plan.blocks.data(), (uint32_t)plan.blocks.size(), plan.streams.data(),
                                        (uint32_t)plan.streams.size(), &flt, &n_rec)))
            die_gpu(api, rc, "ingest_bgzf");
        std::vector<pomfret_gpu_sliced_record> sl(n_rec ? n_rec : 1);
        if ((rc = api.batch_ingest_records(batch, sl.data(), n_rec))) die_gpu(api, rc, "ingest_records");
        const int64_t beg0 = start > 0 ? (int64_t)start - 1 : 0;
        std::vector<pomfret_gpu_read_desc> descs;
        std::vector<uint8_t> read_hap;
        for (uint32_t i = 0; i < n_rec; i++) {
            const pomfret_gpu_sliced_record &S = sl[i];
            if (S.bad) { fprintf(stderr, "[E::%s] malformed alignment record in %s\n", "pomfret", bam.fn.c_str()); exit(1); }
            if (!S.keep || !((int64_t)S.pos < (int64_t)end && (int64_t)S.end_pos > beg0)) continue;
            char buf[256];
            const char *qn = S.qname;
            if (S.l_qname > sizeof(S.qname)) {
                if ((rc = api.batch_ingest_qname(batch, i, buf, sizeof(buf)))) die_gpu(api, rc, "ingest_qname");
                qn = buf;
            }
            auto it = ps.qname2haptag.find(qn);
            if (it == ps.qname2haptag.end()) continue;
            int hap_raw;
            if (ps.stores_raw_tag) {
                auto ir = ps.qname2haptag_raw.find(qn);
                if (ir == ps.qname2haptag_raw.end()) continue;
                hap_raw = ir->second;
            } else {
                hap_raw = S.hp;
                if (S.hp_irregular) fprintf(stderr, "[W::%s] irregular HP tag? qn=%s qs=%d\n", "get_hp_from_aln", qn, (int)S.pos);
            }
            if ((uint8_t)hap_raw == (uint8_t)kHaptagUnphased) continue;
            if (!S.md) die_gpu(api, POMFRET_GPU_ERR_MISSING_MD, "recover");
            pomfret_gpu_read_desc d;
            memset(&d, 0, sizeof(d));
            d.pos = S.pos; d.l_qseq = S.l_qseq; d.n_cigar = S.n_cigar; d.flag = S.flag; d.mapq = S.mapq;
            d.hp = kHaptagUnphased; d.mn = -1; d.ml_len = -1;
            d.cigar = (const uint32_t *)(uintptr_t)S.cigar; d.seq = (const uint8_t *)(uintptr_t)S.seq;
            d.md = (const char *)(uintptr_t)S.md; d.md_len = S.md_len;
            d.reserved = S.end_pos;
            descs.push_back(d);
            read_hap.push_back((uint8_t)it->second);
        }
        std::vector<int32_t> votes(2 * poss.size() + 1, 0);
        if (!descs.empty()) {
            if ((rc = api.batch_add_reads_device(batch, descs.data(), (uint32_t)descs.size(), nullptr))) die_gpu(api, rc, "batch_add_reads");
            if ((rc = api.batch_submit(batch))) die_gpu(api, rc, "batch_submit");
            if ((rc = api.haptag(batch, nullptr, 0, nullptr, 0, nullptr))) die_gpu(api, rc, "haptag");
            std::vector<uint8_t> tags(descs.size());
            std::vector<int32_t> st(descs.size());
            if ((rc = api.batch_collect_haptags(batch, tags.data(), st.data()))) die_gpu(api, rc, "collect_haptags");  // (MD errors surface here)
            if ((rc = api.variant_votes(batch, poss.data(), (uint32_t)poss.size(), read_hap.data(), votes.data()))) die_gpu(api, rc, "variant_votes");
        }
        // the reference's merged walk (blockjoin.c:2557-2600): known positions in order; of several known variants at
        // one position only the last one sees the reads' variants; a known variant that ends the merged list (no read
        // variant at or behind it) is never evaluated
        const size_t n = poss.size();
        for (size_t i = 0; i < n; i++) {
            const bool is_last_entry = i + 1 == n && votes[2 * n] == 0;
            if (is_last_entry) break;
            int c0 = votes[2 * i], c1 = votes[2 * i + 1];
            if (i + 1 < n && poss[i + 1] == poss[i]) c0 = c1 = 0;
            (*pos2hap)[poss[i]] = c0 > c1 ? 1u : c1 > c0 ? 0u : (uint32_t)kHaptagUnphased;
        }
    }

    // estimate_read_coverage_dirtyfast (blockjoin.c:951-1040) for one contig through the compressed ingest: slices of
    // 4 Mb of reference, a record belongs to the slice its start lies in; inflate, record walk, the estimator's own
    // filters (:1000-1011) and the bin increments (:1016-1021) run on the device, the host only adds up one number per slice.
    int coverage_contig_device(int tid) {
        const GpuApi &api = eng->api;
        need_batch();
        if (fd < 0) {
            fd = ::open(bam.fn.c_str(), O_RDONLY);
            struct stat st;
            if (fd < 0 || fstat(fd, &st) != 0) { fprintf(stderr, "[E::%s] failed to open input file: %s\n", "estimate_read_coverage_dirtyfast", bam.fn.c_str()); exit(1); }
            file_size = (uint64_t)st.st_size;
        }
        const uint32_t mod = 5000;
        const int64_t contig_len = (int64_t)bam.hdr->target_len[tid], slice = 4000000;
        const uint32_t n_bins = (uint32_t)(contig_len / mod);
        pomfret_gpu_ingest_filter flt;
        memset(&flt, 0, sizeof(flt));
        flt.min_mapq = 5; flt.min_len = 15000; flt.check_de = 1; flt.max_de = kMinAlnDe;
        uint64_t tot = 0;
        for (int64_t beg = 0; beg < contig_len; beg += slice) {
            const int64_t end = std::min(contig_len, beg + slice);
            plan.clear();
            if (!ingest_plan_region(bam.idx, tid, beg, end, 0, file_size, &plan)) break;  // (no index data for this target)
            if (plan.ranges.empty()) continue;
            int rc;
            if ((rc = api.batch_reset(batch))) die_gpu(api, rc, "batch_reset");
            if (comp_buf.size() < plan.comp_bytes + 64) comp_buf.resize(plan.comp_bytes + plan.comp_bytes / 4 + 64);
            read_chunk(&plan, comp_buf.data());
            uint32_t n_rec = 0;
            if ((rc = api.batch_ingest_bgzf(batch, comp_buf.data(), plan.comp_bytes, plan.blocks.data(), (uint32_t)plan.blocks.size(), plan.streams.data(),
                                            (uint32_t)plan.streams.size(), &flt, &n_rec)))
                die_gpu(api, rc, "ingest_bgzf");
            stats.n_ingest_bytes += plan.comp_bytes;
            uint64_t inc = 0;
            if ((rc = api.batch_ingest_coverage(batch, (uint32_t)beg, mod, n_bins, &inc))) die_gpu(api, rc, "ingest_coverage");
            tot += inc;
        }
        return n_bins ? (int)(tot / n_bins) : 0;
    }

    // output_modify_bam + sam_index_build3 (blockjoin.c:3022-3103, 4714-4731) with the device doing the read side: the
    // file is walked in chunks cut at record starts (the linear index holds them); inflate, record walk and slicing run
    // on the device, the host decides every record's tag from its name (retag_next), retag_kernel lays the records out
    // as the uncompressed output stream with their HP tags set, the host cuts that stream into BGZF blocks exactly as
    // bgzf_write / bgzf_flush_try would (a record that does not fit the open block starts a new one), compresses the
    // blocks on `n_threads` threads with the shim's own block compressor (same bytes as the sequential writer) and
    // builds the BAI from the records' positions and the blocks' addresses instead of reading the output back.
    // Returns false — nothing written that matters, the caller runs the host writer — when the file holds something
    // the raw-record path does not reproduce (CIGARs carried in the CG tag are re-encoded by bam_read1 / bam_write1,
    // HP values outside 1..255).
    bool rewrite_bam_device(const PhaseState &ps, const std::string &fn_out, const std::string &fn_bai, int n_threads) {
        const GpuApi &api = eng->api;
        need_batch();
        if (fd < 0) {
            fd = ::open(bam.fn.c_str(), O_RDONLY);
            struct stat st;
            if (fd < 0 || fstat(fd, &st) != 0) return false;
            file_size = (uint64_t)st.st_size;
        }
        // ---- cut points: the first record, then record starts from the linear index, then the end of the file ----
        const uint64_t first_voff = (uint64_t)bgzf_tell(bam.fp->fp.bgzf);
        std::vector<uint64_t> cuts;
        for (int t = 0; t < pomfret_idx_nref(bam.idx); t++) {
            const uint64_t *lin = nullptr;
            const int n = pomfret_idx_linear(bam.idx, t, &lin);
            for (int i = 0; i < n; i++) if (lin[i] > first_voff) cuts.push_back(lin[i]);
        }
        std::sort(cuts.begin(), cuts.end());
        cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        const uint64_t end_voff = file_size << 16;
        size_t chunk_bytes = (size_t)48 << 20;
        if (const char *e = getenv("POMFRET_REWRITE_CHUNK_MB")) chunk_bytes = (size_t)std::max(1, atoi(e)) << 20;
        std::vector<uint64_t> bounds{first_voff};
        for (uint64_t c : cuts)
            if ((c >> 16) - (bounds.back() >> 16) >= chunk_bytes && c < end_voff) bounds.push_back(c);
        bounds.push_back(end_voff);

        FILE *fo = fopen(fn_out.c_str(), "wb");
        if (!fo) { fprintf(stderr, "[E::%s] failed to open output file: %s\n", "output_modify_bam", fn_out.c_str()); return false; }
        setvbuf(fo, nullptr, _IOFBF, 4 << 20);
        uint64_t file_addr = 0;
        std::vector<uint8_t> cbuf;
        std::vector<int> csize;
        // compress blocks[i] = [ptr, len) in parallel, append them to the file in order; returns each block's address
        auto emit_blocks = [&](const std::vector<std::pair<const uint8_t *, int>> &blocks, std::vector<uint64_t> *addrs) {
            const size_t nb = blocks.size();
            if (cbuf.size() < nb * (size_t)BGZF_MAX_BLOCK_SIZE) cbuf.resize(nb * (size_t)BGZF_MAX_BLOCK_SIZE);
            csize.assign(nb, 0);
            std::atomic<size_t> next(0);
            auto work = [&] {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= nb) break;
                    csize[i] = pomfret_bgzf_compress_block(cbuf.data() + i * (size_t)BGZF_MAX_BLOCK_SIZE, blocks[i].first, blocks[i].second, -1);
                }
            };
            std::vector<std::thread> th;
            const int nt = (int)std::min<size_t>((size_t)std::max(1, n_threads), nb);
            for (int t = 1; t < nt; t++) th.emplace_back(work);
            work();
            for (auto &t : th) t.join();
            for (size_t i = 0; i < nb; i++) {
                if (csize[i] < 0 || fwrite(cbuf.data() + i * (size_t)BGZF_MAX_BLOCK_SIZE, 1, (size_t)csize[i], fo) != (size_t)csize[i]) {
                    fprintf(stderr, "[E::%s] failed to write %s\n", "output_modify_bam", fn_out.c_str());
                    exit(1);
                }
                if (addrs) addrs->push_back(file_addr);
                file_addr += (uint64_t)csize[i];
            }
        };
        // ---- header (bam_hdr_write: its own blocks, flushed) ----
        {
            std::vector<uint8_t> h;
            auto p32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) h.push_back((uint8_t)(v >> (8 * i))); };
            h.insert(h.end(), {'B', 'A', 'M', 1});
            p32((uint32_t)bam.hdr->l_text);
            h.insert(h.end(), bam.hdr->text, bam.hdr->text + bam.hdr->l_text);
            p32((uint32_t)bam.hdr->n_targets);
            for (int i = 0; i < bam.hdr->n_targets; i++) {
                const uint32_t l = (uint32_t)strlen(bam.hdr->target_name[i]) + 1;
                p32(l);
                h.insert(h.end(), bam.hdr->target_name[i], bam.hdr->target_name[i] + l);
                p32(bam.hdr->target_len[i]);
            }
            std::vector<std::pair<const uint8_t *, int>> blocks;
            for (size_t o = 0; o < h.size(); o += BGZF_BLOCK_SIZE) blocks.emplace_back(h.data() + o, (int)std::min<size_t>(BGZF_BLOCK_SIZE, h.size() - o));
            emit_blocks(blocks, nullptr);
        }
        // ---- records ----
        struct RecMeta { int32_t tid; uint32_t pos, end; uint8_t unmapped; uint64_t p0, p1; };
        std::vector<RecMeta> metas;                      // for the index
        std::vector<std::pair<uint64_t, uint64_t>> blk;  // (stream offset of the block's first byte, file address)
        std::vector<uint8_t> ustream;                    // pending bytes of the open block + the chunk's stream
        size_t pending = 0;                              // bytes of the open block carried from the previous chunk
        uint64_t stream_pos = 0;                         // stream offset of ustream[pending]
        RetagCursor cur;
        std::vector<pomfret_gpu_sliced_record> sl;
        std::vector<uint64_t> dst_off;
        std::vector<uint8_t> hp_val;
        bool ok = true;
        pomfret_gpu_ingest_filter flt;
        memset(&flt, 0, sizeof(flt));
        flt.keep_all_flags = 1;
        for (size_t c = 0; ok && c + 1 < bounds.size(); c++) {
            const uint64_t vbeg = bounds[c], vend = bounds[c + 1];
            plan.clear();
            IngestPlan::Range r;
            r.file_off = vbeg >> 16;
            uint64_t stop = (vend >> 16) + ((vend & 0xffff) ? 65536 : 0);
            if (stop > file_size) stop = file_size;
            if (stop <= r.file_off) continue;
            r.bytes = stop - r.file_off; r.comp_off = 0; r.vbeg = vbeg; r.vend = vend; r.run = 0; r.tid = POMFRET_GPU_ANY_TID; r.end0 = 0;
            plan.comp_bytes = r.bytes;
            plan.ranges.push_back(r);
            int rc;
            if ((rc = api.batch_reset(batch))) die_gpu(api, rc, "batch_reset");
            if (comp_buf.size() < plan.comp_bytes + 64) comp_buf.resize(plan.comp_bytes + plan.comp_bytes / 4 + 64);
            read_chunk(&plan, comp_buf.data());
            uint32_t n_rec = 0;
            if ((rc = api.batch_ingest_bgzf(batch, comp_buf.data(), plan.comp_bytes, plan.blocks.data(), (uint32_t)plan.blocks.size(), plan.streams.data(),
                                            (uint32_t)plan.streams.size(), &flt, &n_rec)))
                die_gpu(api, rc, "ingest_bgzf");
            stats.n_ingest_bytes += plan.comp_bytes;
            if (n_rec == 0) continue;
            sl.resize(n_rec);
            if ((rc = api.batch_ingest_records(batch, sl.data(), n_rec))) die_gpu(api, rc, "ingest_records");
            dst_off.resize(n_rec); hp_val.resize(n_rec);
            uint64_t out_bytes = 0;
            const size_t meta0 = metas.size();
            for (uint32_t i = 0; i < n_rec && ok; i++) {
                const pomfret_gpu_sliced_record &S = sl[i];
                if (S.bad) { fprintf(stderr, "[E::%s] malformed alignment record in %s\n", "pomfret", bam.fn.c_str()); exit(1); }
                if (S.cg_cigar) { ok = false; break; }
                char buf[256];
                const char *qn = S.qname;
                if (S.l_qname > sizeof(S.qname)) {
                    if ((rc = api.batch_ingest_qname(batch, i, buf, sizeof(buf)))) die_gpu(api, rc, "ingest_qname");
                    qn = buf;
                }
                const int tid = S.tid;
                const char *refname = tid >= 0 && tid < bam.hdr->n_targets ? bam.hdr->target_name[tid] : "";
                if (!ps.stores_raw_tag && S.hp_irregular) fprintf(stderr, "[W::%s] irregular HP tag? qn=%s qs=%d\n", "get_hp_from_aln", qn, (int)S.pos);
                const int hp = retag_next(ps, &cur, tid, refname, qn, (int)S.pos, S.hp);
                const int val = hp + 1;
                if (val < 1 || val > 255) { ok = false; break; }
                hp_val[i] = (uint8_t)val;
                // the record's size behind bam_aux_update_int (kernel: retag_kernel)
                const uint32_t sz = val < 255 ? 1u : 2u;
                const uint32_t t = S.hp_type;
                const uint32_t old_sz = !t ? 0u : (t == 'c' || t == 'C') ? 1u : (t == 's' || t == 'S') ? 2u : (t == 'i' || t == 'I') ? 4u : 0u;
                uint32_t grow = 0;
                if (!t) grow = 3u + sz;
                else if (old_sz == 0u) grow = 0;          // not an integer tag: the update fails, the record stays as it is
                else if (old_sz < sz) grow = sz - old_sz;
                dst_off[i] = out_bytes;
                const uint64_t nbytes = (uint64_t)S.rec_bytes + grow;
                metas.push_back({tid, S.pos, S.end_pos, (uint8_t)((S.flag & 4u) != 0), stream_pos + out_bytes, stream_pos + out_bytes + nbytes});
                out_bytes += nbytes;
            }
            if (!ok) break;
            if (ustream.size() < pending + out_bytes + 64) ustream.resize(pending + out_bytes + out_bytes / 8 + 64);
            if ((rc = api.batch_ingest_retag(batch, dst_off.data(), hp_val.data(), out_bytes, ustream.data() + pending))) die_gpu(api, rc, "ingest_retag");
            // ---- cut into blocks (bgzf_flush_try before a record, bgzf_write's flush when a block is full) ----
            std::vector<std::pair<const uint8_t *, int>> blocks;
            std::vector<uint64_t> starts;
            size_t open_at = 0;              // offset in ustream of the open block's first byte
            const uint64_t base_pos = stream_pos - pending;  // stream offset of ustream[0]
            auto cut = [&](size_t at) {
                if (at > open_at) { blocks.emplace_back(ustream.data() + open_at, (int)(at - open_at)); starts.push_back(base_pos + open_at); }
                open_at = at;
            };
            size_t at = pending;
            for (size_t i = meta0; i < metas.size(); i++) {
                size_t rem = (size_t)(metas[i].p1 - metas[i].p0);
                if ((at - open_at) + rem > (size_t)BGZF_BLOCK_SIZE) cut(at);
                while (rem) {
                    const size_t room = (size_t)BGZF_BLOCK_SIZE - (at - open_at);
                    const size_t n = std::min(rem, room);
                    at += n; rem -= n;
                    if (at - open_at == (size_t)BGZF_BLOCK_SIZE) cut(at);
                }
            }
            std::vector<uint64_t> addrs;
            emit_blocks(blocks, &addrs);
            for (size_t i = 0; i < addrs.size(); i++) blk.emplace_back(starts[i], addrs[i]);
            // the open block's bytes stay for the next chunk
            const size_t tail = at - open_at;
            memmove(ustream.data(), ustream.data() + open_at, tail);
            pending = tail;
            stream_pos += out_bytes;
        }
        if (!ok) { fclose(fo); return false; }
        if (pending) {
            std::vector<std::pair<const uint8_t *, int>> blocks{{ustream.data(), (int)pending}};
            std::vector<uint64_t> addrs;
            emit_blocks(blocks, &addrs);
            blk.emplace_back(stream_pos - pending, addrs[0]);
        }
        const uint64_t eof_addr = file_addr;
        {
            const uint8_t *e = nullptr;
            const int n = pomfret_bgzf_eof_block(&e);
            if (fwrite(e, 1, (size_t)n, fo) != (size_t)n) { fprintf(stderr, "[E::%s] failed to write %s\n", "output_modify_bam", fn_out.c_str()); exit(1); }
        }
        if (fclose(fo) != 0) { fprintf(stderr, "[E::%s] failed to write %s\n", "output_modify_bam", fn_out.c_str()); exit(1); }
        fprintf(stderr, "[M::%s] bam written. now indexing...\n", "main_blockjoin");
        // ---- index (sam_index_build3): virtual offsets from the block table ----
        blk.emplace_back(stream_pos, eof_addr);  // a position at the end of the stream names the block behind the last one
        pomfret_bai_builder *B = pomfret_bai_new(bam.hdr->n_targets);
        size_t k = 0;
        auto voff = [&](uint64_t p) {
            while (k + 1 < blk.size() && blk[k + 1].first <= p) k++;
            return (blk[k].second << 16) | (p - blk[k].first);
        };
        int stat = 0;
        for (const RecMeta &m : metas) {
            const uint64_t off0 = voff(m.p0), off1 = voff(m.p1);
            if (pomfret_bai_add(B, m.tid, (hts_pos_t)(int32_t)m.pos, (hts_pos_t)m.end, m.unmapped, off0, off1) != 0) { stat = -1; break; }
        }
        const int rs = pomfret_bai_finish(B, stat == 0 ? fn_bai.c_str() : nullptr);
        if (stat == 0) stat = rs;
        if (stat != 0) fprintf(stderr, "[W::%s] failed to build index for output bam (status code=%d)\n", "main_blockjoin", stat);
        fprintf(stderr, "[M::%s] bam index written.\n", "main_blockjoin");
        return true;
    }

    // pre_haplotagging_read_in_one_ref through the compressed ingest: the contig is walked in slices of 2 Mb of
    // reference; a slice's query returns every record that overlaps it, and a record belongs to the slice its start
    // lies in, so every record is taken once, in BAM order.  The device inflates, slices and haplotags; the host sees
    // one header per record.
    void haptag_contig_device(const std::string &chrom, const KnownVariants &kv, TagMap *raw) {
        const GpuApi &api = eng->api;
        double t0 = now_s();
        need_batch();
        const int tid = sam_hdr_name2tid(bam.hdr, chrom.c_str());
        if (tid < 0) return;
        if (fd < 0) {
            fd = ::open(bam.fn.c_str(), O_RDONLY);
            struct stat st;
            if (fd < 0 || fstat(fd, &st) != 0) { fprintf(stderr, "[E::%s] failed to open input file: %s\n", "pre_haplotagging_read_in_one_ref", bam.fn.c_str()); exit(1); }
            file_size = (uint64_t)st.st_size;
        }
        const int64_t contig_len = (int64_t)bam.hdr->target_len[tid], slice = 2000000;
        uint32_t prev_i_left = 0;
        int n_new[4] = {0, 0, 0, 0};
        pomfret_gpu_ingest_filter flt;
        memset(&flt, 0, sizeof(flt));  // primary records only (blockjoin.c:1862)
        std::vector<pomfret_gpu_sliced_record> sl;
        std::vector<pomfret_gpu_read_desc> descs;
        std::vector<uint32_t> known_first, which;
        for (int64_t beg = 0; beg < contig_len; beg += slice) {
            const int64_t end = std::min(contig_len, beg + slice);
            plan.clear();
            if (!ingest_plan_region(bam.idx, tid, beg, end, 0, file_size, &plan)) { fprintf(stderr, "[E::%s] region query failed for %s\n", "pre_haplotagging_read_in_one_ref", chrom.c_str()); exit(1); }
            if (plan.ranges.empty()) continue;
            int rc;
            if ((rc = api.batch_reset(batch))) die_gpu(api, rc, "batch_reset");
            if (comp_buf.size() < plan.comp_bytes + 64) comp_buf.resize(plan.comp_bytes + plan.comp_bytes / 4 + 64);
            uint8_t *comp = comp_buf.data();
            std::string err;
            if (!ingest_read(fd, &plan, comp, &err)) { fprintf(stderr, "[E::%s] %s: %s\n", "pomfret", bam.fn.c_str(), err.c_str()); exit(1); }
            uint32_t n_rec = 0;
            if ((rc = api.batch_ingest_bgzf(batch, comp, plan.comp_bytes, plan.blocks.data(), (uint32_t)plan.blocks.size(), plan.streams.data(),
                                            (uint32_t)plan.streams.size(), &flt, &n_rec)))
                die_gpu(api, rc, "ingest_bgzf");
            sl.resize(n_rec ? n_rec : 1);
            if ((rc = api.batch_ingest_records(batch, sl.data(), n_rec))) die_gpu(api, rc, "ingest_records");
            stats.n_ingest_bytes += plan.comp_bytes;
            descs.clear(); known_first.clear(); which.clear();
            for (uint32_t i = 0; i < n_rec; i++) {
                const pomfret_gpu_sliced_record &S = sl[i];
                if (S.bad) { fprintf(stderr, "[E::%s] malformed alignment record in %s\n", "pomfret", bam.fn.c_str()); exit(1); }
                if (!S.keep || (int64_t)S.pos < beg) continue;  // (a record that starts in an earlier slice was taken there)
                if (!S.md) die_gpu(api, POMFRET_GPU_ERR_MISSING_MD, "haptag");
                which.push_back(i);
                stats.n_haptag_reads++;
                stats.n_haptag_bases += S.l_qseq;
                if (kv.vars.empty()) continue;
                uint32_t k = prev_i_left;  // i_left cursor, blockjoin.c:1716-1720
                while (k < kv.vars.size() && kv.vars[k].pos < S.pos) k++;
                prev_i_left = k == 0 ? 0 : k - 1;
                known_first.push_back(k);
                pomfret_gpu_read_desc d;
                memset(&d, 0, sizeof(d));
                d.pos = S.pos; d.l_qseq = S.l_qseq; d.n_cigar = S.n_cigar; d.flag = S.flag; d.mapq = S.mapq;
                d.hp = kHaptagUnphased; d.mn = -1; d.ml_len = -1;
                d.cigar = (const uint32_t *)(uintptr_t)S.cigar; d.seq = (const uint8_t *)(uintptr_t)S.seq;
                d.md = (const char *)(uintptr_t)S.md; d.md_len = S.md_len;
                d.reserved = S.end_pos;
                descs.push_back(d);
            }
            std::vector<uint8_t> tags(which.size() ? which.size() : 1, (uint8_t)kHaptagUnphased);
            if (!kv.vars.empty() && !descs.empty()) {
                if ((rc = api.batch_add_reads_device(batch, descs.data(), (uint32_t)descs.size(), nullptr))) die_gpu(api, rc, "batch_add_reads");
                if ((rc = api.batch_submit(batch))) die_gpu(api, rc, "batch_submit");
                if ((rc = api.haptag(batch, kv.vars.data(), (uint32_t)kv.vars.size(), kv.bases.data(), (uint32_t)kv.bases.size(), known_first.data())))
                    die_gpu(api, rc, "haptag");
                std::vector<int32_t> st(descs.size());
                if ((rc = api.batch_collect_haptags(batch, tags.data(), st.data()))) die_gpu(api, rc, "collect_haptags");
            }
            for (size_t j = 0; j < which.size(); j++) {
                const pomfret_gpu_sliced_record &S = sl[which[j]];
                char buf[256];
                const char *qn = S.qname;
                if (S.l_qname > sizeof(S.qname)) {
                    if ((rc = api.batch_ingest_qname(batch, which[j], buf, sizeof(buf)))) die_gpu(api, rc, "ingest_qname");
                    qn = buf;
                }
                auto ins = raw->emplace(qn, (int)tags[j]);  // first alignment wins (blockjoin.c:1880-1889)
                if (ins.second) n_new[tags[j] == 0 ? 0 : tags[j] == 1 ? 1 : 2]++; else n_new[3]++;
            }
        }
        fprintf(stderr, "[dbg::%s] tagged: %d new hap0, %d new hap1, %d new unphased, %d dup\n", "pre_haplotagging_read_in_one_ref",
                n_new[0], n_new[1], n_new[2], n_new[3]);
        stats.t_haptag += now_s() - t0;
    }

    // pre_haplotagging_read_in_one_ref (blockjoin.c:1841-1898): every primary record of the contig, in BAM order
    // into `raw` (first alignment of a name wins, :1880-1889)
    void haptag_contig(const std::string &chrom, const KnownVariants &kv, TagMap *raw) {
        if (eng->gpu_ingest) { haptag_contig_device(chrom, kv, raw); return; }
        const GpuApi &api = eng->api;
        double t0 = now_s();
        need_batch();
        hts_itr_t *itr = sam_itr_querys(bam.idx, bam.hdr, chrom.c_str());
        if (!itr) return;
        std::vector<const char *> names;
        std::vector<uint32_t> known_first;
        std::vector<pomfret_gpu_read_desc> descs;
        uint32_t prev_i_left = 0;
        size_t bytes = 0;
        int n_new[4] = {0, 0, 0, 0};
        arena.rewind();
        auto flush = [&]() {
            if (names.empty()) return;
            std::vector<uint8_t> tags(names.size(), (uint8_t)kHaptagUnphased);
            if (!kv.vars.empty()) {
                int rc = api.batch_reset(batch);
                if (rc) die_gpu(api, rc, "batch_reset");
                if ((rc = api.batch_add_reads(batch, descs.data(), (uint32_t)descs.size()))) die_gpu(api, rc, "batch_add_reads");
                if ((rc = api.batch_submit(batch))) die_gpu(api, rc, "batch_submit");
                if ((rc = api.haptag(batch, kv.vars.data(), (uint32_t)kv.vars.size(), kv.bases.data(), (uint32_t)kv.bases.size(), known_first.data())))
                    die_gpu(api, rc, "haptag");
                std::vector<int32_t> st(names.size());
                if ((rc = api.batch_collect_haptags(batch, tags.data(), st.data()))) die_gpu(api, rc, "collect_haptags");
            }
            for (size_t i = 0; i < names.size(); i++) {
                auto ins = raw->emplace(names[i], (int)tags[i]);  // first alignment wins (blockjoin.c:1880-1889)
                if (ins.second) n_new[tags[i] == 0 ? 0 : tags[i] == 1 ? 1 : 2]++; else n_new[3]++;
            }
            names.clear(); known_first.clear(); descs.clear(); bytes = 0;
            arena.rewind();
        };
        while (next_record(itr)) {
            const bam1_t *b = &view;
            const int flag = b->core.flag;
            if ((flag & 4) || (flag & 256) || (flag & 2048)) continue;
            if (!bam_aux_get(b, "MD")) die_gpu(api, POMFRET_GPU_ERR_MISSING_MD, "haptag");
            names.push_back(bam_get_qname(b));
            arena.commit((size_t)b->l_data);
            stats.n_haptag_reads++;
            stats.n_haptag_bases += (uint64_t)b->core.l_qseq;
            if (!kv.vars.empty()) {
                // i_left cursor, blockjoin.c:1716-1720
                uint32_t i = prev_i_left;
                const uint32_t start_pos = (uint32_t)b->core.pos;
                while (i < kv.vars.size() && kv.vars[i].pos < start_pos) i++;
                prev_i_left = i == 0 ? 0 : i - 1;
                known_first.push_back(i);
                pomfret_gpu_read_desc d;
                describe_record(b, kHaptagUnphased, &d);
                d.mm = nullptr; d.mm_len = 0; d.ml = nullptr; d.ml_len = -1;  // the haplotagger needs CIGAR, SEQ and MD only
                descs.push_back(d);
            }
            bytes += (size_t)b->l_data;
            if (names.size() >= 32768 || bytes >= ((size_t)768 << 20)) flush();
        }
        flush();
        hts_itr_destroy(itr);
        fprintf(stderr, "[dbg::%s] tagged: %d new hap0, %d new hap1, %d new unphased, %d dup\n", "pre_haplotagging_read_in_one_ref",
                n_new[0], n_new[1], n_new[2], n_new[3]);
        stats.t_haptag += now_s() - t0;
    }
};

pomfret_gpu_config base_config(const Options &o) {
    pomfret_gpu_config c;
    memset(&c, 0, sizeof(c));
    c.lo = o.lo; c.hi = o.hi; c.min_mapq = o.mapq; c.k = o.k; c.k_span = o.k_span;
    c.cov_known = o.cov; c.cov_for_selection = o.cov_for_selection;
    c.cov_for_runtime = c.cov_for_selection * 2;  // blockjoin.c:4657
    c.readlen_threshold = o.readlen_threshold;
    c.n_candidates_per_iter = o.n_candidates_per_iter;
    return c;
}

bool files_exist(const Options &o) {  // sancheck_cliopt_t_files_exist, blockjoin.c:4606-4641
    htsFile *fp = hts_open(o.fn_bam.c_str(), "rb");
    if (!fp || !fp->is_bgzf) { fprintf(stderr, "[E::%s] cannot open bam file: %s\n", "sancheck_cliopt_t_files_exist", o.fn_bam.c_str()); if (fp) hts_close(fp); return false; }
    hts_close(fp);
    for (const std::string *f : {&o.fn_vcf, &o.fn_tsv, &o.fn_gtf}) {
        if (f->empty()) continue;
        FILE *t = fopen(f->c_str(), "r");
        if (!t) { fprintf(stderr, "[E::%s] cannot open %s\n", "sancheck_cliopt_t_files_exist", f->c_str()); return false; }
        fclose(t);
    }
    return true;
}

void check_limits(const pomfret_gpu_config &c) {
    if (c.k > 8 || c.n_candidates_per_iter > 1024) {
        fprintf(stderr, "[E::%s] this build supports methmer k <= 8 and <= 1024 candidates per iteration (got k=%d, n=%d)\n", "pomfret",
                c.k, c.n_candidates_per_iter);
        exit(1);
    }
}

// Run all windows of all contigs.  The ordered (contig, window) list is cut into chunks of consecutive windows
// (one batch each) and the chunk list into one contiguous region set per device, balanced by the estimated
// number of records; a worker serves its device's set first and helps the others when it
// runs dry.  Nothing is exchanged between devices: the host gathers one decision per window and one tag per read.
void run_windows(Engine &eng, const Options &opt, const PhaseState &ps, const std::vector<pomfret_gpu_config> &cfg_per_ref,
                 std::vector<std::vector<WindowOut>> *results, RunStats *stats) {
    struct Chunk { int i_ref; std::vector<WindowJob> jobs; uint64_t cost; };
    std::vector<Chunk> chunks;
    // windows per batch: the option is the upper bound; a run with few windows is cut finer so that every worker
    // has something to inflate (the batches are then small, but a batch costs microseconds of launches)
    size_t n_windows_total = 0;
    for (const Ranges &rg : ps.st.ranges) n_windows_total += rg.n;
    int per = opt.windows_per_batch > 0 ? opt.windows_per_batch : 8;
    {
        size_t feeders = (size_t)std::max(1, opt.threads);
        if (eng.gpu_ingest) feeders = std::min<size_t>(feeders, 3 * (size_t)std::max(1, eng.expected_devices(opt.gpus)));
        per = (int)std::max<size_t>(1, std::min<size_t>((size_t)per, n_windows_total / (2 * feeders) + 1));
    }
    results->assign(ps.st.ref_names.size(), {});
    uint64_t total_cost = 0;
    for (size_t r = 0; r < ps.st.ref_names.size(); r++) {
        const Ranges &rg = ps.st.ranges[r];
        (*results)[r].assign(rg.n, WindowOut());
        for (size_t i = 0; i < rg.n; i += (size_t)per) {
            Chunk c;
            c.i_ref = (int)r;
            c.cost = 0;
            for (size_t j = i; j < rg.n && j < i + (size_t)per; j++) {
                c.jobs.push_back({(int)r, j, rg.starts[j], rg.ends[j]});
                c.cost += (uint64_t)(rg.ends[j] - rg.starts[j]) + 2 * kReadback;  // ~ records fetched for the window
            }
            total_cost += c.cost;
            chunks.push_back(std::move(c));
        }
    }
    // (the engine may still be starting: the region sets are cut for the expected number of devices, and a worker's
    //  device index is taken modulo the real count once it is known)
    const int n_dev = std::max(1, eng.expected_devices(opt.gpus));
    std::vector<size_t> set_begin((size_t)n_dev + 1, chunks.size());
    {
        uint64_t acc = 0;
        size_t c = 0;
        for (int d = 0; d < n_dev; d++) {
            set_begin[(size_t)d] = c;
            const uint64_t target = total_cost * (uint64_t)(d + 1) / (uint64_t)n_dev;
            while (c < chunks.size() && (acc < target || d == n_dev - 1)) acc += chunks[c++].cost;
        }
        set_begin[(size_t)n_dev] = chunks.size();
    }
    std::vector<std::atomic<size_t>> cursor((size_t)n_dev);
    for (int d = 0; d < n_dev; d++) cursor[(size_t)d].store(set_begin[(size_t)d]);
    // With the compressed ingest a worker only reads file ranges and sorts record headers: three of them keep a device
    // busy, and more only queue up behind the driver's allocation locks.  The host loader inflates on the workers
    // themselves and takes every thread it is given.
    int n_workers = std::max(1, std::min<int>(opt.threads, (int)chunks.size()));
    if (eng.gpu_ingest) n_workers = std::min(n_workers, 3 * n_dev);
    if (const char *e = getenv("POMFRET_WORKERS")) n_workers = std::max(1, atoi(e));
    std::mutex mu;
    auto body = [&](int wid) {
        Worker wk;
        wk.eng = &eng; wk.id = wid; wk.device = wid % n_dev;
        const double tw0 = now_s();
        if (!wk.open(opt.fn_bam)) exit(1);
        const double tw1 = now_s();
        size_t n_own = 0, n_helped = 0, n_ahead = 0;
        auto process = [&](size_t c, bool own) {
            const Chunk &ch = chunks[c];
            std::vector<WindowOut> outs;
            wk.run_chunk(ps.st.ref_names[ch.i_ref], ch.jobs, cfg_per_ref[ch.i_ref], ps.stores_raw_tag ? &ps.qname2haptag_raw : nullptr, &outs);
            for (size_t j = 0; j < ch.jobs.size(); j++) (*results)[ch.i_ref][ch.jobs[j].i_win] = std::move(outs[j]);
            (own ? n_own : n_helped)++;
        };
        // while the CUDA start-up is still under way: read the blocks of this worker's next chunks into plain memory
        // (POMFRET_READ_AHEAD_MB per worker, default 96: enough for the first chunks to be ready when the device is; a
        //  page-fault storm over gigabytes of fresh buffers would hold up the driver's own start-up, which maps memory
        //  in this same process), so that the file reads are done by the time the device can take them
        if (eng.gpu_ingest) {
            std::deque<Worker::ReadAhead> queue;
            size_t buffered = 0;
            const size_t d = (size_t)wk.device;
            size_t ahead_cap = (size_t)96 << 20;
            if (const char *e = getenv("POMFRET_READ_AHEAD_MB")) ahead_cap = (size_t)std::max(0, atoi(e)) << 20;
            while (!eng.is_ready() && buffered < ahead_cap) {
                const size_t c = cursor[d].fetch_add(1);
                if (c >= set_begin[d + 1]) break;
                queue.emplace_back();
                Worker::ReadAhead &ra = queue.back();
                ra.chunk = c;
                wk.plan_chunk(ps.st.ref_names[chunks[c].i_ref], chunks[c].jobs, &ra.plan);
                ra.comp.resize(ra.plan.comp_bytes + 64);
                wk.read_chunk(&ra.plan, ra.comp.data());
                buffered += ra.comp.size();
            }
            n_ahead = queue.size();
            for (Worker::ReadAhead &ra : queue) {
                wk.ahead = &ra;
                process(ra.chunk, true);
                wk.ahead = nullptr;
                std::vector<uint8_t>().swap(ra.comp);
            }
        }
        for (int k = 0; k < n_dev; k++) {
            const size_t d = (size_t)((wk.device + k) % n_dev);  // own region set first, then the others'
            for (;;) {
                const size_t c = cursor[d].fetch_add(1);
                if (c >= set_begin[d + 1]) break;
                process(c, k == 0);
            }
        }
        (void)n_ahead;
        const double tw2 = now_s();
        wk.close();
        fprintf(stderr, "[T::worker %d] phases: plan+read %.3f, wait engine/batch %.3f, reset %.3f, (read) %.3f, ingest %.3f, records d2h %.3f, assign %.3f | submit+decode %.3f, ..pileup %.3f (cumulative)\n",
                wid, wk.stats.t_phase[0], wk.stats.t_phase[1], wk.stats.t_phase[2], wk.stats.t_phase[3], wk.stats.t_phase[4], wk.stats.t_phase[5], wk.stats.t_phase[6],
                wk.stats.t_phase[7], wk.stats.t_phase[8]);
        fprintf(stderr, "[T::worker %d] device %d: %zu chunks of its region set, %zu of others; open %.2fs, chunks %.2fs (load %.2fs, gpu %.2fs), close %.2fs\n",
                wid, wk.device, n_own, n_helped, tw1 - tw0, tw2 - tw1, wk.stats.t_load, wk.stats.t_gpu, now_s() - tw2);
        std::lock_guard<std::mutex> lock(mu);
        stats->n_windows += wk.stats.n_windows; stats->n_reads += wk.stats.n_reads; stats->n_bases += wk.stats.n_bases;
        stats->n_shared += wk.stats.n_shared; stats->n_ingest_bytes += wk.stats.n_ingest_bytes;
        stats->t_load += wk.stats.t_load; stats->t_gpu += wk.stats.t_gpu;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_workers; t++) th.emplace_back(body, t);
    body(0);
    for (auto &t : th) t.join();
}

// load_intervals_from_file with the -u pre-pass hooked in (blockjoin.c:4446-4468).  The reference tags one
// contig after the other on the main thread; here the contigs are handed to workers (one device each) as the
// VCF reader finishes them, every contig into its own table, and the tables are merged in contig order with
// first-insert-wins (blockjoin.c:1880-1889) — what the sequential run produces.
bool load_all_intervals(Engine &eng, const Options &opt, PhaseState *ps, RunStats *stats) {
    std::string fatal;
    const std::string fn_interval = !opt.fn_tsv.empty() ? opt.fn_tsv : !opt.fn_gtf.empty() ? opt.fn_gtf : opt.fn_vcf;
    const IntervalFormat fmt = !opt.fn_tsv.empty() ? IntervalFormat::TSV : !opt.fn_gtf.empty() ? IntervalFormat::GTF : IntervalFormat::VCF;
    if (opt.bam_needs_haplotagging) {
        struct Task { std::string chrom; KnownVariants kv; TagMap tags; };
        std::deque<Task> tasks;  // stable addresses
        std::mutex mu;
        std::condition_variable cv;
        size_t next_task = 0;
        bool done_reading = false;
        int n_workers = std::max(1, opt.threads);
        if (eng.gpu_ingest) n_workers = std::min(n_workers, 3 * std::max(1, eng.expected_devices(opt.gpus)));
        std::vector<std::thread> th;
        auto body = [&](int wid) {
            Worker wk;
            wk.eng = &eng; wk.id = wid; wk.device = wid;  // (taken modulo the device count when the batch is created)
            bool opened = false;
            for (;;) {
                Task *t = nullptr;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return next_task < tasks.size() || done_reading; });
                    if (next_task >= tasks.size()) break;
                    t = &tasks[next_task++];
                }
                if (!opened) { if (!wk.open(opt.fn_bam)) exit(1); opened = true; }
                wk.haptag_contig(t->chrom, t->kv, &t->tags);
            }
            if (opened) wk.close();
            std::lock_guard<std::mutex> lk(mu);
            stats->n_haptag_reads += wk.stats.n_haptag_reads; stats->n_haptag_bases += wk.stats.n_haptag_bases; stats->t_haptag += wk.stats.t_haptag;
        };
        for (int t = 0; t < n_workers; t++) th.emplace_back(body, t);
        ps->stores_raw_tag = true;
        bool ok = load_intervals(opt.fn_vcf, IntervalFormat::VCF, &ps->st,
                                 [&](const std::string &chrom, KnownVariants &kv, bool) {
                                     { std::lock_guard<std::mutex> lk(mu); tasks.push_back(Task{chrom, kv, TagMap()}); }
                                     cv.notify_one();
                                 },
                                 &fatal);
        { std::lock_guard<std::mutex> lk(mu); done_reading = true; }
        cv.notify_all();
        for (auto &t : th) t.join();
        for (Task &t : tasks)
            for (auto &kv : t.tags) ps->qname2haptag_raw.emplace(kv.first, kv.second);
        if (!ok) { fprintf(stderr, "[E::%s] failed to open file for phase blocks: %s\n", "load_intervals_from_file", opt.fn_vcf.c_str()); exit(1); }
        if (!fatal.empty()) { fprintf(stderr, "%s\n", fatal.c_str()); exit(1); }
        size_t n_loaded = 0;
        for (const Ranges &r : ps->st.ranges) n_loaded += r.n;
        if (n_loaded == 0) {
            fprintf(stderr, "[E::%s] Nothing loaded from vcf (ref_n=%d), cannot haptag the input bam. Terminating.\n", "blockjoin_parallel", (int)ps->st.ref_names.size());
            exit(1);
        }
        if (fmt != IntervalFormat::VCF) {  // gtf/tsv override the vcf's phase blocks
            ps->st = Storage();
            if (!load_intervals(fn_interval, fmt, &ps->st, nullptr, &fatal)) { fprintf(stderr, "[E::%s] failed to open file for phase blocks: %s\n", "load_intervals_from_file", fn_interval.c_str()); exit(1); }
        }
    } else {
        if (!load_intervals(fn_interval, fmt, &ps->st, nullptr, &fatal)) { fprintf(stderr, "[E::%s] failed to open file for phase blocks: %s\n", "load_intervals_from_file", fn_interval.c_str()); exit(1); }
        if (!fatal.empty()) { fprintf(stderr, "%s\n", fatal.c_str()); exit(1); }
    }
    return true;
}

}  // namespace

std::vector<int> estimate_read_coverage(const std::string &fn_bam) {
    BamReader bam;
    std::vector<int> covs;
    if (!bam.open(fn_bam)) return covs;
    covs.assign((size_t)bam.hdr->n_targets, 0);
    hts_itr_t *itr = sam_itr_querys(bam.idx, bam.hdr, ".");
    fprintf(stderr, "[M::%s] estimate read depths...\n", "estimate_read_coverage_dirtyfast");
    const int mod = 5000;
    std::vector<uint64_t> buf;
    int prev = -1, refID = -1;
    auto close_ref = [&](int id) {
        if (buf.empty()) { covs[id] = 0; return; }
        uint64_t tot = 0;
        for (uint64_t v : buf) tot += v;
        covs[id] = (int)(tot / buf.size());
    };
    bam1_t *b = bam.rec;
    while (sam_itr_next(bam.fp, itr, b) >= 0) {
        refID = b->core.tid;
        if (refID < 0) continue;
        if (refID > bam.hdr->n_targets) continue;
        if (refID != prev) {
            if (prev >= 0) close_ref(prev);
            buf.assign((size_t)(bam.hdr->target_len[refID] / mod), 0);
            prev = refID;
        }
        const int flag = b->core.flag;
        if ((flag & 4) || (flag & 256) || (flag & 2048)) continue;
        if (b->core.qual < 5) continue;
        float de = -1;
        uint8_t *t = bam_aux_get(b, "de");
        if (t) de = (float)bam_aux2f(t);
        if ((uint32_t)b->core.l_qseq < 15000) continue;
        if (de > kMinAlnDe) continue;
        const uint32_t s = (uint32_t)b->core.pos, e = (uint32_t)bam_endpos(b);
        for (int i = (int)s; i < (int)e; i += mod) {
            size_t bin = (size_t)(i / mod);
            if (bin < buf.size()) buf[bin]++;  // the reference can write one bin past the end here
        }
    }
    if (refID >= 0) close_ref(refID);
    hts_itr_destroy(itr);
    for (int i = 0; i < bam.hdr->n_targets; i++)
        fprintf(stderr, "[M::%s] %s est. coverage is %d\n", "estimate_read_coverage_dirtyfast", bam.hdr->target_name[i], covs[(size_t)i]);
    return covs;
}

// The same on the device: contigs are handed to the feeder workers, see Worker::coverage_contig_device
std::vector<int> estimate_read_coverage_device(Engine &eng, const Options &opt) {
    const double T = now_s();
    BamReader hdr;
    std::vector<int> covs;
    if (!hdr.open(opt.fn_bam)) return covs;
    const int n_targets = hdr.hdr->n_targets;
    covs.assign((size_t)n_targets, 0);
    fprintf(stderr, "[M::%s] estimate read depths...\n", "estimate_read_coverage_dirtyfast");
    std::atomic<int> next(0);
    int n_workers = std::max(1, std::min(opt.threads, n_targets));
    n_workers = std::min(n_workers, 3 * std::max(1, eng.expected_devices(opt.gpus)));
    auto body = [&](int wid) {
        Worker wk;
        wk.eng = &eng; wk.id = wid; wk.device = wid;
        bool opened = false;
        for (;;) {
            const int tid = next.fetch_add(1);
            if (tid >= n_targets) break;
            if (!opened) { if (!wk.open(opt.fn_bam)) exit(1); opened = true; }
            covs[(size_t)tid] = wk.coverage_contig_device(tid);
        }
        if (opened) wk.close();
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_workers; t++) th.emplace_back(body, t);
    body(0);
    for (auto &t : th) t.join();
    for (int i = 0; i < n_targets; i++)
        fprintf(stderr, "[M::%s] %s est. coverage is %d\n", "estimate_read_coverage_dirtyfast", hdr.hdr->target_name[i], covs[(size_t)i]);
    fprintf(stderr, "[T::%s] used %.1fs\n", "estimate_read_coverage_dirtyfast", now_s() - T);
    return covs;
}

// --gpus absent: as many of the visible devices as the input can keep busy — one per 4 GiB of BAM file.  Every device
// costs a CUDA context (0.3 s and more of start-up each) and its feeder workers their arenas; a small input is done
// on one device before a second one would be ready.  An explicit --gpus N is taken as it is.
Options resolve_gpus(const Options &in) {
    Options o = in;
    if (o.gpus <= 0) {
        struct stat st;
        if (stat(o.fn_bam.c_str(), &st) == 0) o.gpus = (int)std::min<uint64_t>(1024, std::max<uint64_t>(1, ((uint64_t)st.st_size + ((uint64_t)4 << 30) - 1) / ((uint64_t)4 << 30)));
    }
    return o;
}

int run_methphase(const Options &opt_in, RunStats *stats) {
    const Options opt = resolve_gpus(opt_in);
    const double T = now_s();
    if (!files_exist(opt)) return 1;
    Engine eng;
    eng.start_async(opt.gpus, opt.threads);
    PhaseState ps;
    if (!load_all_intervals(eng, opt, &ps, stats)) return 1;
    size_t n_loaded = 0;
    for (const Ranges &r : ps.st.ranges) n_loaded += r.n;
    if (n_loaded == 0) { fprintf(stderr, "[E::%s] No intervals loaded, terminating.\n", "blockjoin_parallel"); exit(1); }
    fprintf(stderr, "[M::%s] input has %d references\n", "blockjoin_parallel", (int)ps.st.ref_names.size());
    for (Ranges &r : ps.st.ranges) { store_raw_intervals(&r); merge_close_intervals(&r, kReadback); }
    fprintf(stderr, "[M::%s] loaded phase block gaps.\n\n\n", "blockjoin_parallel");

    // per-contig parameters (blockjoin_one_chrom_callback, blockjoin.c:4357-4392)
    pomfret_gpu_config base = base_config(opt);
    std::vector<int> covs;
    BamReader hdr_only;
    if (base.cov_for_selection <= 0) {
        covs = eng.gpu_ingest ? estimate_read_coverage_device(eng, opt) : estimate_read_coverage(opt.fn_bam);
        if (!hdr_only.open(opt.fn_bam)) return 1;
    }
    std::vector<pomfret_gpu_config> cfgs;
    for (size_t r = 0; r < ps.st.ref_names.size(); r++) {
        pomfret_gpu_config c = base;
        if (c.cov_for_selection <= 0) {
            int ref_i = sam_hdr_name2tid(hdr_only.hdr, ps.st.ref_names[r].c_str());
            if (ref_i < 0) { fprintf(stderr, "pomfret: blockjoin_one_chrom_callback: Assertion `ref_i>=0' failed.\n"); abort(); }
            const int coverage = covs[(size_t)ref_i];
            c.cov_for_selection = coverage / 10 + 1;
            c.cov_for_runtime = c.cov_for_selection * 2;
            c.n_candidates_per_iter = coverage / 4 + 1;
        }
        if (c.cov_for_selection <= 0) { fprintf(stderr, "[W::%s] had to clamp cov_for_selection (ref: %s)\n", "blockjoin_one_chrom_callback", ps.st.ref_names[r].c_str()); c.cov_for_selection = 1; }
        if (c.n_candidates_per_iter <= 1) { fprintf(stderr, "[W::%s] had to clamp n_candidates_per_iter (ref: %s)\n", "blockjoin_one_chrom_callback", ps.st.ref_names[r].c_str()); c.n_candidates_per_iter = 2; }
        fprintf(stderr, "[dbg::%s] ref %s using: cov_for_selection=%d, n_cand_per_iter=%d\n", "blockjoin_one_chrom_callback", ps.st.ref_names[r].c_str(), c.cov_for_selection, c.n_candidates_per_iter);
        check_limits(c);
        cfgs.push_back(c);
    }
    std::vector<std::vector<WindowOut>> results;
    fprintf(stderr, "[T::%s] setup + intervals %.2fs\n", "run_methphase", now_s() - T);
    run_windows(eng, opt, ps, cfgs, &results, stats);
    fprintf(stderr, "[T::%s] windows done at %.2fs\n", "run_methphase", now_s() - T);
    // decisions + per-contig tag tables, then the global table in contig order (first insert wins both times)
    for (size_t r = 0; r < ps.st.ref_names.size(); r++) {
        Ranges &rg = ps.st.ranges[r];
        TagMap local;
        for (size_t i = 0; i < rg.n; i++) {
            rg.decisions[i] = results[r][i].decision;
            for (auto &kv : results[r][i].tags) local.emplace(kv.first, kv.second);
        }
        for (auto &kv : local) ps.qname2haptag.emplace(kv.first, kv.second);
    }
    fprintf(stderr, "\n\n[M::%s] done, used %.1fs.\n", "blockjoin_parallel", now_s() - T);

    lift_decisions(&ps.st);
    make_flips_onraw(&ps.st);
    generate_new_phase_blocks(&ps.st);
    if (opt.write_debug_files) output_debug_read2tag(ps, opt.output_prefix);
    output_gtf(ps, opt.output_prefix);
    fprintf(stderr, "[M::%s] gtf written.\n", "main_blockjoin");
    if (opt.do_output_tsv) { output_tsv(ps, opt.output_prefix); fprintf(stderr, "[M::%s] tsv written.\n", "main_blockjoin"); }
    if (!opt.fn_vcf.empty()) {
        fprintf(stderr, "[M::%s] writing vcf...\n", "main_blockjoin");
        if (eng.gpu_ingest) {
            Worker wk;
            wk.eng = &eng; wk.id = 0; wk.device = 0;
            if (!wk.open(opt.fn_bam)) exit(1);
            recover_variant_phase_in_dropped_intervals(&ps, opt.fn_bam, opt.fn_vcf,
                [&](const std::string &refname, uint32_t start, uint32_t end, const std::vector<uint32_t> &poss, std::unordered_map<uint32_t, uint32_t> *pos2hap) {
                    wk.recover_interval_device(ps, refname, start, end, poss, pos2hap);
                });
            wk.close();
        } else recover_variant_phase_in_dropped_intervals(&ps, opt.fn_bam, opt.fn_vcf);
        output_modify_vcf(opt.fn_vcf, ps, opt.output_prefix);
        fprintf(stderr, "[M::%s] vcf written.\n", "main_blockjoin");
    }
    fprintf(stderr, "[T::%s] gtf/vcf written at %.2fs\n", "run_methphase", now_s() - T);
    if (opt.do_output_bam) {
        const std::string fn_bam_out = opt.output_prefix + ".mp.bam", fn_bai_out = opt.output_prefix + ".mp.bam.bai";
        bool done = false;
        if (eng.gpu_ingest && !getenv("POMFRET_HOST_BAM_REWRITE")) {
            Worker wk;
            wk.eng = &eng; wk.id = 0; wk.device = 0;
            if (!wk.open(opt.fn_bam)) exit(1);
            done = wk.rewrite_bam_device(ps, fn_bam_out, fn_bai_out, std::max(opt.threads, opt.threads_bam));
            wk.close();
            if (!done) fprintf(stderr, "[M::%s] records the device re-tagger does not copy verbatim (CG-tag CIGARs): host writer\n", "output_modify_bam");
        }
        if (!done) {
            output_modify_bam(opt.fn_bam, ps, fn_bam_out);
            fprintf(stderr, "[M::%s] bam written. now indexing...\n", "main_blockjoin");
            int stat = sam_index_build3(fn_bam_out.c_str(), fn_bai_out.c_str(), 0, opt.threads_bam);
            if (stat != 0) fprintf(stderr, "[W::%s] failed to build index for output bam (status code=%d)\n", "main_blockjoin", stat);
            fprintf(stderr, "[M::%s] bam index written.\n", "main_blockjoin");
        }
    }
    stats->t_total = now_s() - T;
    return 0;
}

int run_report(const Options &opt_in, RunStats *stats) {
    const Options opt = resolve_gpus(opt_in);
    const double T = now_s();
    if (opt.fn_bam.empty()) { fprintf(stderr, "[E::%s] input bam file name missing\n", "main_methreport"); exit(1); }
    if (opt.fn_vcf.empty()) { fprintf(stderr, "[E::%s] input vcf file name missing\n", "main_methreport"); exit(1); }
    {
        BamReader probe;
        if (!probe.open(opt.fn_bam)) { fprintf(stderr, "[E::%s] failed to open input bam: %s\n", "main_methreport", opt.fn_bam.c_str()); exit(1); }
    }
    const std::string fn_out = opt.output_prefix + ".report.tsv";
    FILE *fp_out = fopen(fn_out.c_str(), "w");
    if (!fp_out) { fprintf(stderr, "[E::%s] failed to open output file\n", "main_methreport"); exit(1); }
    Engine eng;
    eng.start_async(opt.gpus, opt.threads);
    PhaseState ps;
    Options o2 = opt;
    o2.fn_tsv.clear(); o2.fn_gtf.clear();  // report always derives its blocks from the vcf (blockjoin.c:4958)
    if (!load_all_intervals(eng, o2, &ps, stats)) return 1;
    // replace the gaps by windows inside the phased stretches (blockjoin.c:4961-4993)
    for (size_t r = 0; r < ps.st.ref_names.size(); r++) {
        Ranges &rg = ps.st.ranges[r];
        std::vector<uint32_t> starts, ends;
        uint32_t prev = rg.abs_start;
        for (size_t i = 0; i < rg.n; i++) {
            const uint32_t start = rg.starts[i], end = rg.ends[i];
            if ((uint32_t)(start - prev) > (uint32_t)opt.chunk_size)
                for (uint32_t p = prev; p + (uint32_t)opt.chunk_stride < start; p += (uint32_t)opt.chunk_stride) {
                    starts.push_back(p);
                    ends.push_back(p + (uint32_t)opt.chunk_size);
                }
            prev = end;
        }
        rg.starts = starts; rg.ends = ends; rg.n = starts.size();
        rg.decisions.assign(starts.size(), -1); rg.n_decisions = starts.size();
        fprintf(stderr, "[M::%s] %s has %d intervals\n", "main_methreport", ps.st.ref_names[r].c_str(), (int)starts.size());
    }
    const int read_coverage = opt.cov;
    std::vector<int> covs;
    BamReader hdr_only;
    if (read_coverage <= 0) {
        fprintf(stderr, "[M::%s] estimating read depths..\n", "main_methreport");
        covs = eng.gpu_ingest ? estimate_read_coverage_device(eng, opt) : estimate_read_coverage(opt.fn_bam);
        if (!hdr_only.open(opt.fn_bam)) return 1;
    }
    std::vector<pomfret_gpu_config> cfgs;
    for (size_t r = 0; r < ps.st.ref_names.size(); r++) {
        pomfret_gpu_config c = base_config(opt);
        // the reference indexes covs[] by contig order of the vcf here (blockjoin.c:5045-5051)
        const int cov = read_coverage <= 0 ? (r < covs.size() ? covs[r] : 0) : read_coverage;
        c.cov_for_selection = cov / 10 + 1;
        c.cov_for_runtime = c.cov_for_selection * 2;
        c.n_candidates_per_iter = cov / 4 + 1;
        check_limits(c);
        cfgs.push_back(c);
    }
    std::vector<std::vector<WindowOut>> results;
    run_windows(eng, opt, ps, cfgs, &results, stats);
    float n_switch = 0, n_fail = 0, n_correct = 0;
    int tot = 0;
    for (size_t r = 0; r < ps.st.ref_names.size(); r++) {
        const Ranges &rg = ps.st.ranges[r];
        for (size_t i = 0; i < rg.n; i++) {
            const int start = (int)rg.starts[i], end = (int)rg.ends[i], decision = results[r][i].decision;
            fprintf(fp_out, "%s\t%d\t%d\t", ps.st.ref_names[r].c_str(), start, end);
            if (decision == 0) { n_correct++; fprintf(fp_out, "correct\n"); }
            else if (decision == 1) { n_switch++; fprintf(fp_out, "switch\n"); }
            else { n_fail++; fprintf(fp_out, "fail\n"); }
            tot++;
            if (tot % 100 == 0)
                fprintf(stdout, "Parsed N=%d regions, currently at %s:%d-%d, correct/(correct+switch)=%.2f%%, correct/N=%.2f%%\n", tot,
                        ps.st.ref_names[r].c_str(), start, end, n_correct / (n_correct + n_switch) * 100.0, n_correct / (float)tot * 100.0);
        }
    }
    fprintf(stdout, "Total N=%d regions, correct/(correct+switch)=%.2f%%, correct/N=%.2f%%\n", tot,
            n_correct / (n_correct + n_switch) * 100.0, n_correct / (float)tot * 100.0);
    fprintf(stderr, "[M::%s] Total N=%d regions, correct/(correct+switch)=%.2f%%, correct/N=%.2f%%\n", "main_methreport", tot,
            n_correct / (n_correct + n_switch) * 100.0, n_correct / (float)tot * 100.0);
    fclose(fp_out);
    fprintf(stderr, "[M::%s] done, used %.1fs\n", "main_methreport", now_s() - T);
    stats->t_total = now_s() - T;
    return 0;
}

}  // namespace pomfret
