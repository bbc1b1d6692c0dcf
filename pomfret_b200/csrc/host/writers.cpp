// Output writers, see writers.h.
#include "writers.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <zlib.h>

namespace pomfret {

namespace {

struct LineReader {  // complete ('\n' terminated) lines only, like the reference's gzread loops
    gzFile fp = nullptr;
    std::string pending;
    char buf[1 << 16];
    bool open(const std::string &fn) { fp = gzopen(fn.c_str(), "rb"); return fp != nullptr; }
    ~LineReader() { if (fp) gzclose(fp); }
    bool next(std::string *line) {
        for (;;) {
            size_t nl = pending.find('\n');
            if (nl != std::string::npos) { line->assign(pending, 0, nl); pending.erase(0, nl + 1); return true; }
            int n = gzread(fp, buf, sizeof(buf));
            if (n <= 0) return false;
            pending.append(buf, (size_t)n);
        }
    }
};

int ref_index(const Storage &st, const std::string &name) {
    for (size_t i = 0; i < st.ref_names.size(); i++) if (st.ref_names[i] == name) return (int)i;
    return -1;
}

// search_substr_idx on a length-limited field
int field_index_n(const char *s, int n, const char *q) {
    const int lq = (int)strlen(q);
    int col = 0, start = 0;
    for (int i = 0; i <= n; i++) {
        if (i == n || s[i] == ':') {
            if (i - start == lq && strncmp(q, s + start, (size_t)lq) == 0) return col;
            if (i == n) break;
            start = i + 1;
            col++;
        }
    }
    return -1;
}
bool field_by_index(const char *s, int idx, int *start, int *len) {
    int col = 0, b = 0;
    for (int i = 0;; i++) {
        if (s[i] == ':' || s[i] == 0) {
            if (col == idx) { *start = b; *len = i - b; return true; }
            if (s[i] == 0) break;
            b = i + 1;
            col++;
        }
    }
    return false;
}

int posint_digits(int n) {
    int d = 1;
    while (n >= 10 && d < 10) { n /= 10; d++; }
    return d;
}

}  // namespace

void output_tsv(const PhaseState &ps, const std::string &prefix) {
    const std::string fn = prefix + ".mp.tsv";
    FILE *fp = fopen(fn.c_str(), "w");
    if (!fp) { fprintf(stderr, "[E::%s] failed to open output file: %s\n", "output_tsv", fn.c_str()); exit(1); }
    int n_blocks = 0;
    for (size_t r = 0; r < ps.st.ref_names.size(); r++)
        for (const U32Pair &b : ps.st.ranges[r].phaseblocks) {
            fprintf(fp, "%s\t%d\t%d\n", ps.st.ref_names[r].c_str(), (int)b.s, (int)b.e);
            n_blocks++;
        }
    fclose(fp);
    fprintf(stderr, "[M::%s] wrote tsv (%d refs, total %d blocks)\n", "output_tsv", (int)ps.st.ref_names.size(), n_blocks);
}

void output_gtf(const PhaseState &ps, const std::string &prefix) {
    const std::string fn = prefix + ".mp.gtf";
    FILE *fp = fopen(fn.c_str(), "w");
    if (!fp) { fprintf(stderr, "[E::%s] failed to open output file: %s\n", "output_gtf", fn.c_str()); exit(1); }
    int n_blocks = 0;
    for (size_t r = 0; r < ps.st.ref_names.size(); r++)
        for (const U32Pair &b : ps.st.ranges[r].phaseblocks) {
            const int start = (int)b.s, end = (int)b.e;
            if (start == 0 || end == 0) continue;  // placeholders (blockjoin.c:2743)
            fprintf(fp, "%s\tPhasing\texon\t%d\t%d\t.\t+\t.\tgene_id \"%d\"; transcript_id \"%d.1\"\n",
                    ps.st.ref_names[r].c_str(), start, end, start, start);
            n_blocks++;
        }
    fclose(fp);
    fprintf(stderr, "[M::%s] wrote gtf (%d refs, total %d blocks)\n", "output_gtf", (int)ps.st.ref_names.size(), n_blocks);
}

void output_debug_read2tag(const PhaseState &ps, const std::string &prefix) {
    const std::string fn = prefix + ".mp.dbg.read2tag";
    FILE *fp = fopen(fn.c_str(), "w");
    if (!fp) { fprintf(stderr, "[E::%s] failed to open output file: %s\n", "dbgoutput_intermediate_read_haplotags", fn.c_str()); exit(1); }
    std::vector<std::pair<std::string, int>> v(ps.qname2haptag.begin(), ps.qname2haptag.end());
    std::sort(v.begin(), v.end());  // the reference writes in hash-table order; sorted here (debug file only)
    for (auto &kv : v) fprintf(fp, "%s\t-1\t%d\n", kv.first.c_str(), (kv.second < 0 ? kHaptagUnphased : kv.second) + 1);
    fclose(fp);
}

// ---------------- read variants on the host (positions only) ----------------

int host_parse_read_variants(const bam1_t *b, std::vector<HostReadVariant> *out) {
    const uint32_t *cigar = bam_get_cigar(b);
    uint32_t ref_pos = (uint32_t)b->core.pos;
    for (uint32_t i = 0; i < b->core.n_cigar; i++) {
        uint32_t c;
        memcpy(&c, cigar + i, 4);
        const uint32_t op = c & 15u, L = c >> 4;
        if (op == 3 || op == 2 || op == 0 || op == 7 || op == 8) ref_pos += L;
        else if (op == 1) out->push_back({ref_pos, L, 2, 0});
    }
    uint8_t *tag = bam_aux_get(b, "MD");
    if (!tag) return POMFRET_GPU_ERR_MISSING_MD;  // assert(tagd), blockjoin.c:1596
    const char *md = bam_aux2Z(tag);
    if (!md || !md[0]) return md ? 0 : POMFRET_GPU_ERR_BAD_MD;
    auto cls = [](int ch) {
        if (ch >= '0' && ch <= '9') return 0;
        if (ch == '^') return 1;
        switch (ch) {
        case 'A': case 'C': case 'G': case 'T': case 'U': case 'N':
        case 'a': case 'c': case 'g': case 't': case 'u': case 'n': return 2;
        default: return 4;
        }
    };
    ref_pos = (uint32_t)b->core.pos;
    int prev = cls((unsigned char)md[0]);
    size_t prev_i = 0;
    if (prev == 2) { out->push_back({ref_pos, 1, 1, 0}); ref_pos++; prev = -1; }
    if (prev >= 4) return POMFRET_GPU_ERR_BAD_MD;
    for (size_t i = 1; md[i]; i++) {
        const int t = cls((unsigned char)md[i]);
        if (t == 4) return POMFRET_GPU_ERR_BAD_MD;
        if (t == prev) continue;
        if (prev == 0) {
            int l = 0;
            for (size_t j = prev_i; j < i; j++) l = l * 10 + (md[j] - '0');
            ref_pos += (uint32_t)l;
        } else if (prev == 1) {
            if (t == 0) {
                const uint32_t L = (uint32_t)(i - prev_i - 1);
                out->push_back({ref_pos, L, 3, 0});
                ref_pos += L;
                prev = t;
                prev_i = i;
            }
            continue;
        }
        if (t == 2) { out->push_back({ref_pos, 1, 1, 0}); ref_pos++; prev = -1; prev_i = i; }
        else { prev = t; prev_i = i; }
    }
    return 0;
}

// ---------------- dropped-interval rescue ----------------

static void recover_one_interval(PhaseState *ps, BamReader &bam, const std::string &refname, uint32_t start, uint32_t end,
                                 const std::vector<uint32_t> &poss, std::unordered_map<uint32_t, uint32_t> *pos2hap) {
    std::vector<uint64_t> pb;
    for (uint32_t p : poss) pb.push_back(((uint64_t)p) << 33);
    std::vector<HostReadVariant> rv;
    char region[1024];
    snprintf(region, sizeof(region), "%s:%d-%d", refname.c_str(), (int)start, (int)end);
    hts_itr_t *itr = sam_itr_querys(bam.idx, bam.hdr, region);
    bam1_t *b = bam.rec;
    while (itr && sam_itr_next(bam.fp, itr, b) >= 0) {
        const char *qn = bam_get_qname(b);
        auto it = ps->qname2haptag.find(qn);
        if (it == ps->qname2haptag.end()) continue;
        const uint8_t hap_meth = (uint8_t)it->second;
        uint8_t hap_raw;
        if (ps->stores_raw_tag) {
            auto ir = ps->qname2haptag_raw.find(qn);
            if (ir == ps->qname2haptag_raw.end()) continue;
            hap_raw = (uint8_t)ir->second;
        } else hap_raw = (uint8_t)hp_from_record(b);
        if (hap_raw == kHaptagUnphased) continue;
        const size_t before = rv.size();
        int rc = host_parse_read_variants(b, &rv);
        if (rc == POMFRET_GPU_ERR_MISSING_MD) { fprintf(stderr, "pomfret: blockjoin.c:1596: parse_variants_for_one_read: Assertion `tagd' failed.\n"); abort(); }
        if (rc == POMFRET_GPU_ERR_BAD_MD) { fprintf(stderr, "[E::%s] invalid MD\n", "parse_variants_for_one_read"); exit(1); }
        for (size_t i = before; i < rv.size(); i++) rv[i].haptag = (uint8_t)(hap_meth << 4 | hap_raw);
        // (the reference's coverage bump compares packed keys with positions and never fires for pos > 0)
    }
    if (itr) hts_itr_destroy(itr);
    if (pb.empty()) return;
    const uint64_t typebit = 1ull << 32;
    for (uint32_t i = 0; i < rv.size(); i++) pb.push_back(((uint64_t)rv[i].pos) << 33 | typebit | i);
    std::sort(pb.begin(), pb.end());
    for (size_t i = 0; i + 1 < pb.size();) {  // i < pb.n-1: a trailing known variant is never evaluated
        if (pb[i] & typebit) { i++; continue; }
        const uint32_t ref_pos = (uint32_t)(pb[i] >> 33);
        int cnt[2] = {0, 0};
        size_t j;
        for (j = i + 1; j < pb.size(); j++) {
            if (!(pb[j] & typebit)) break;
            if ((uint32_t)(pb[j] >> 33) != ref_pos) break;
            const int hap = rv[(uint32_t)pb[j]].haptag >> 4;
            if (hap == 0 || hap == 1) cnt[hap]++;
        }
        uint32_t hap_of_ref = cnt[0] > cnt[1] ? 1u : cnt[1] > cnt[0] ? 0u : (uint32_t)kHaptagUnphased;
        (*pos2hap)[ref_pos] = hap_of_ref;
        i = j;
    }
}

int recover_variant_phase_in_dropped_intervals(PhaseState *ps, const std::string &fn_bam, const std::string &fn_vcf,
                                               const RecoverIntervalFn &on_device) {
    const size_t n_ref = ps->st.ref_names.size();
    // known variants per contig: a second pass over the VCF keyed by contig name (blockjoin.c:2626-2640)
    std::vector<KnownVariants> vars(n_ref);
    {
        Storage st2;
        std::string fatal;
        load_intervals(fn_vcf, IntervalFormat::VCF, &st2,
                       [&](const std::string &chrom, KnownVariants &kv, bool) {
                           int i = ref_index(ps->st, chrom);
                           if (i >= 0) { vars[i].vars.insert(vars[i].vars.end(), kv.vars.begin(), kv.vars.end()); }
                       }, &fatal);
    }
    ps->varphase_in_dropped.assign(n_ref, {});
    BamReader bam;
    bool opened = false;
    std::vector<uint32_t> poss;
    for (size_t r = 0; r < n_ref; r++) {
        const Ranges &rg = ps->st.ranges[r];
        size_t prev_i = 0;
        for (const U32Pair &d : rg.dropped) {
            const uint32_t start = d.s - 1, end = d.e + 1;
            poss.clear();
            for (size_t i = prev_i; i < vars[r].vars.size(); i++) {
                const uint32_t pos = vars[r].vars[i].pos;
                if (pos >= start && pos < end) poss.push_back(pos);
                if (pos >= end) { prev_i = i; break; }
            }
            if (on_device) { on_device(ps->st.ref_names[r], start, end, poss, &ps->varphase_in_dropped[r]); continue; }
            if (!opened) {
                if (!bam.open(fn_bam)) { fprintf(stderr, "[E::%s] failed to open input file: %s\n", "recover_variant_phase_in_one_interval", fn_bam.c_str()); exit(1); }
                opened = true;
            }
            recover_one_interval(ps, bam, ps->st.ref_names[r], start, end, poss, &ps->varphase_in_dropped[r]);
        }
    }
    return 0;
}

// ---------------- VCF ----------------

namespace {

// get_new_phaseblock_ID1, blockjoin.c:2365-2381
int new_phaseblock_id(const Ranges &r, int pos) {
    for (const U32Pair &b : r.phaseblocks) {
        if (b.s == UINT32_MAX || b.e == 0 || b.e == UINT32_MAX) continue;
        if ((uint32_t)pos >= b.s && (uint32_t)pos < b.e) return (int)b.s;
    }
    return -1;
}

// get_flip_status, blockjoin.c:2438-2473 (prev_idx is the caller's cursor and is only reset when the
// position decreases, exactly like the reference)
int flip_status(const Ranges &rr, int *prev_idx, int pos) {
    const int n = (int)rr.raw.size();
    int j;
    for (j = *prev_idx; j < n; j++) {
        if (j < 0) continue;
        const int start = (int)rr.raw[j].s;
        if (start >= pos) {
            *prev_idx = j == 0 ? 0 : j - 1;
            int stat = *prev_idx < (int)rr.flips_onraw.size() ? rr.flips_onraw[*prev_idx] : 0;
            if (pos <= (int)rr.raw[0].s) stat = 0;
            return stat;
        }
    }
    *prev_idx = j - 1;
    const int last = n == 0 ? 0 : n - 1;
    return last < (int)rr.flips_onraw.size() ? rr.flips_onraw[last] : 0;
}

}  // namespace

int output_modify_vcf(const std::string &fn_vcf, const PhaseState &ps, const std::string &prefix) {
    LineReader rd;
    if (!rd.open(fn_vcf)) { fprintf(stderr, "[E::%s] failed to open input file: %s\n", "output_modify_vcf", fn_vcf.c_str()); exit(1); }
    const std::string fn_out = prefix + ".mp.vcf";
    FILE *out = fopen(fn_out.c_str(), "w");
    if (!out) { fprintf(stderr, "[E::%s] failed to open output file: %s\n", "output_modify_vcf", fn_out.c_str()); exit(1); }
    int n_modified = 0, n_failed = 0, n_tot = 0;
    int last_pos = -1, prev_block_idx = 0;
    std::string line, nl;
    while (rd.next(&line)) {
        n_tot++;
        const char *s = line.c_str();
        const int s_l = (int)line.size();
        int altered = 0;
        do {
            if (s[0] == '#') {
                if (s[1] == '#') break;
                int n = 1;
                for (int i = 0; i < s_l; i++) n += s[i] == '\t';
                if (n < 10) { fprintf(stderr, "[E::%s] vcf only has %d columns; mandatory >=8; we also need FORMAT and at least 1 sample\n", "alter_vcf_line", n); exit(1); }
                if (n > 10) { fprintf(stderr, "[E::%s] multi-sample vcf not implemented, TODO/TBD\n", "alter_vcf_line"); exit(1); }
                break;
            }
            int col = 0, start = 0, pos = 0, i_ps = -1, i_gt = -1, i_ref = 0;
            std::string refname;
            for (int i = 0; i < s_l; i++) {
                if (s[i] != '\t') continue;
                if (col == 0) {
                    refname.assign(s + start, (size_t)(i - start));
                    i_ref = ref_index(ps.st, refname);
                    pos = 0; i_ps = -1; i_gt = -1;
                    if (i_ref < 0) break;
                } else if (col == 1) {
                    pos = atoi(std::string(s + start, (size_t)(i - start)).c_str());
                    if (pos < last_pos) prev_block_idx = 0;  // "we've encountered a new chromosome"
                    last_pos = pos;
                } else if (col == 8) {
                    i_ps = field_index_n(s + start, i - start, "PS");
                    i_gt = field_index_n(s + start, i - start, "GT");
                }
                col++;
                start = i + 1;
            }
            if (pos == 0 || i_ps < 0 || i_gt < 0 || i_ref < 0) break;
            int ps_start, ps_l, gt_start, gt_l;
            if (!field_by_index(s + start, i_ps, &ps_start, &ps_l) || !field_by_index(s + start, i_gt, &gt_start, &gt_l)) break;
            if (ps_l == 1 && s[start + ps_start] == '.') break;
            if (gt_l < 3) break;
            const char *GT = s + start + gt_start;
            if (GT[1] != '|') break;
            if ((GT[0] != '0' && GT[0] != '1') || (GT[2] != '0' && GT[2] != '1')) break;
            const Ranges &rr = ps.st.ranges[i_ref];
            const int groupID = new_phaseblock_id(rr, pos);
            bool is_dropped = false;
            for (const U32Pair &d : rr.dropped) if ((uint32_t)pos >= d.s && (uint32_t)pos <= d.e) { is_dropped = true; break; }
            const int need_flip = flip_status(rr, &prev_block_idx, pos);
            bool is_middle = false;
            if (groupID >= 0 && is_dropped && (size_t)i_ref < ps.varphase_in_dropped.size()) {
                auto it = ps.varphase_in_dropped[i_ref].find((uint32_t)(pos - 1));
                if (it != ps.varphase_in_dropped[i_ref].end() && (it->second == 0 || it->second == 1)) is_middle = true;
            }
            if (groupID < 0 || is_dropped) {
                if (!is_middle) break;
                nl.assign(s, (size_t)(start + ps_start));
                nl += ".";
                nl += s + start + ps_start + ps_l;
                if ((size_t)(start + gt_start + 1) < nl.size()) nl[start + gt_start + 1] = '/';
                altered = 2;
            } else {
                nl.assign(s, (size_t)(start + ps_start));
                char num[16];
                snprintf(num, sizeof(num), "%d", groupID);
                num[posint_digits(groupID)] = 0;
                nl += num;
                nl += s + start + ps_start + ps_l;
                if (need_flip) {
                    const size_t g0 = (size_t)(start + gt_start);
                    if (g0 + 2 < nl.size()) {
                        nl[g0] = nl[g0] == '0' ? '1' : '0';
                        nl[g0 + 2] = nl[g0] == '0' ? '1' : '0';
                    }
                }
                altered = 1;
            }
        } while (0);
        if (!altered) fprintf(out, "%s\n", s);
        else {
            if (altered == 2) n_failed++; else n_modified++;
            fprintf(out, "%s\n", nl.c_str());
        }
    }
    fclose(out);
    fprintf(stderr, "[M::%s] wrote vcf output, (%d ok + %d dropped)/%d lines modified \n", "output_modify_vcf", n_modified, n_failed, n_tot);
    return 0;
}

// ---------------- BAM ----------------

// The tag a record gets in the output BAM (blockjoin.c:3056-3092): stateful over the records in file order
// (check_if_in_phased_intervals keeps a cursor per contig, the flip status changes when the cursor moves).
int retag_next(const PhaseState &ps, RetagCursor *c, int tid, const char *refname, const char *qn, int start_pos, int hp_of_record) {
    if (tid != c->prev_tid) { c->prev_unphased_idx = 1; c->prev_tid = tid; }
    int hp_raw;
    if (ps.stores_raw_tag) {
        auto it = ps.qname2haptag_raw.find(qn);
        hp_raw = it == ps.qname2haptag_raw.end() ? kHaptagUnphased : it->second;
    } else hp_raw = hp_of_record;
    // check_if_in_phased_intervals, blockjoin.c:2406-2426 (merged gap arrays, cursor starts at 1)
    const int i_ref = ref_index(ps.st, refname);
    bool updated = false;
    if (i_ref >= 0) {
        const Ranges &r = ps.st.ranges[i_ref];
        const int prev = c->prev_unphased_idx;
        for (int j = c->prev_unphased_idx; j < (int)r.n; j++) {
            if (j < 1) continue;
            if ((uint32_t)start_pos >= r.ends[j - 1] && (uint32_t)start_pos <= r.starts[j]) {
                if (j != prev) { updated = true; c->prev_unphased_idx = j; }
                break;
            }
        }
        if (updated) {
            const int idx = c->prev_unphased_idx - 1;
            c->need_flip = idx >= 0 && idx < (int)r.flips_onraw.size() ? r.flips_onraw[idx] : 0;
        }
    }
    // get_read_new_haplotag, blockjoin.c:2990-3020
    int hp;
    auto it = ps.qname2haptag.find(qn);
    if (it != ps.qname2haptag.end()) { hp = it->second; if (c->need_flip) hp ^= 1; }
    else { hp = hp_raw; if ((hp == 0 || hp == 1) && c->need_flip) hp ^= 1; }
    return hp;
}

int output_modify_bam(const std::string &fn_bam, const PhaseState &ps, const std::string &fn_out) {
    BamReader in;
    if (!in.open(fn_bam)) { fprintf(stderr, "[E::%s] failed to open input bam: %s\n", "output_modify_bam", fn_bam.c_str()); return 1; }
    hts_itr_t *itr = sam_itr_querys(in.idx, in.hdr, ".");
    BGZF *out = bgzf_open(fn_out.c_str(), "w");
    if (!out) { fprintf(stderr, "[E::%s] failed to open output file: %s\n", "output_modify_bam", fn_out.c_str()); return 1; }
    if (bam_hdr_write(out, in.hdr) != 0) { fprintf(stderr, "[E::%s] failed to write bam header\n", "output_modify_bam"); exit(1); }
    RetagCursor cur;
    bam1_t *b = in.rec;
    while (sam_itr_next(in.fp, itr, b) >= 0) {
        const int tid = b->core.tid;
        const char *refname = tid >= 0 && tid < in.hdr->n_targets ? in.hdr->target_name[tid] : "";
        const char *qn = bam_get_qname(b);
        const int start_pos = (int)b->core.pos;
        const int hp = retag_next(ps, &cur, tid, refname, qn, start_pos, ps.stores_raw_tag ? kHaptagUnphased : hp_from_record(b));
        bam_aux_update_int(b, "HP", hp + 1);
        if (bam_write1(out, b) < 0) fprintf(stderr, "[E::%s] failed to write bam entry (ref=%s pos=%d qn=%s)\n", "output_modify_bam", refname, start_pos, qn);
    }
    bgzf_close(out);
    hts_itr_destroy(itr);
    return 0;
}

}  // namespace pomfret
