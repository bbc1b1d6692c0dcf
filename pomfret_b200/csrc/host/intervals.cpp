// Phase-block bookkeeping, see intervals.h.
#include "intervals.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <zlib.h>

namespace pomfret {

namespace {

// strtok_r(s, "\t") semantics: runs of tabs are one separator, empty fields do not exist
void split_tabs(char *line, std::vector<char *> *out) {
    out->clear();
    char *p = line;
    while (*p) {
        while (*p == '\t') p++;
        if (!*p) break;
        out->push_back(p);
        while (*p && *p != '\t') p++;
        if (*p) *p++ = 0;
    }
}

// search_substr_idx(s, q, ':', get_idx=1): index of the first ':' separated field equal to q, or -1
int field_index(const char *s, const char *q) {
    const size_t lq = strlen(q);
    int col = 0;
    const char *start = s;
    for (const char *p = s;; p++) {
        if (*p == ':' || *p == 0) {
            if ((size_t)(p - start) == lq && strncmp(q, start, lq) == 0) return col;
            if (*p == 0) break;
            start = p + 1;
            col++;
        }
    }
    return -1;
}

// get_substr_by_idx(s, idx, ':'): the idx-th ':' separated field
bool field_by_index(const char *s, int idx, const char **start, int *len) {
    int col = 0;
    const char *b = s;
    for (const char *p = s;; p++) {
        if (*p == ':' || *p == 0) {
            if (col == idx) { *start = b; *len = (int)(p - b); return true; }
            if (*p == 0) break;
            b = p + 1;
            col++;
        }
    }
    return false;
}

uint8_t nt4(char c) {
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3; default: return 4;
    }
}

// Reads '\n' terminated lines from a plain or gz file; a last line without '\n' is not delivered
// (reference readers only act on complete lines, blockjoin.c:2016-2147).
struct LineReader {
    gzFile fp = nullptr;
    std::string pending;
    char buf[1 << 16];
    bool open(const std::string &fn) { fp = gzopen(fn.c_str(), "rb"); return fp != nullptr; }
    ~LineReader() { if (fp) gzclose(fp); }
    bool next(std::string *line) {
        for (;;) {
            size_t nl = pending.find('\n');
            if (nl != std::string::npos) {
                line->assign(pending, 0, nl);
                pending.erase(0, nl + 1);
                return true;
            }
            int n = gzread(fp, buf, sizeof(buf));
            if (n <= 0) return false;
            pending.append(buf, (size_t)n);
        }
    }
};

}  // namespace

bool variant_from_vcf_fields(const char *ref, const char *alt, const char *format, const char *sample, uint32_t pos1,
                             KnownVariants *out) {
    const int i_gt = field_index(format, "GT");
    if (i_gt < 0) return false;
    const char *gt;
    int gt_l;
    if (!field_by_index(sample, i_gt, &gt, &gt_l)) return false;
    if (gt_l != 3 || gt[1] != '|') return false;
    if ((gt[0] != '0' && gt[0] != '1') || (gt[2] != '0' && gt[2] != '1')) return false;
    const int ref_l = (int)strlen(ref), alt_l = (int)strlen(alt);
    uint32_t pos = pos1 - 1;
    pomfret_gpu_variant v;
    memset(&v, 0, sizeof(v));
    const char *chars;
    if (ref_l == 1 && alt_l == 1) { v.op = 1; v.len = 1; chars = alt; }
    else if (ref_l == alt_l) {
        fprintf(stderr, "[W::%s] unhandled variant case at pos %u ref=%s alt=%s\n", "insert_variant_from_vcf_line", pos + 1, ref, alt);
        return false;
    } else if (ref_l > alt_l) { v.op = 3; v.len = (uint32_t)(ref_l - alt_l); pos += 1; chars = ref + 1; }
    else { v.op = 2; v.len = (uint32_t)(alt_l - ref_l); chars = alt + 1; }
    v.pos = pos;
    v.haptag = (uint8_t)(gt[0] - '0');
    v.bases_off = (uint32_t)out->bases.size();
    for (uint32_t i = 0; i < v.len; i++) out->bases.push_back(nt4(chars[i]));
    out->vars.push_back(v);
    return true;
}

bool load_intervals(const std::string &fn, IntervalFormat fmt, Storage *st, const ContigVariantsFn &on_contig,
                    std::string *fatal) {
    LineReader rd;
    if (!rd.open(fn)) return false;
    std::string line;
    std::vector<char *> tok;
    std::vector<char> scratch;
    uint32_t prev_end = UINT32_MAX, prev_group = UINT32_MAX;  // prev_group survives contig changes (App. A.8)
    int cur = -1;  // index of the contig the current line belongs to
    KnownVariants vars;
    const bool want_vars = (bool)on_contig && fmt == IntervalFormat::VCF;
    while (rd.next(&line)) {
        if (!line.empty() && line[0] == '#') continue;
        scratch.assign(line.begin(), line.end());
        scratch.push_back(0);
        split_tabs(scratch.data(), &tok);
        if (tok.empty()) continue;
        const char *chrom = tok[0];
        if (st->ref_names.empty()) {
            st->ref_names.push_back(chrom);
            st->ranges.emplace_back();
            cur = 0;
            fprintf(stderr, "[M::%s] at ref %s\n", "load_intervals_from_file", chrom);
        } else {
            int found = -1;
            for (int i = (int)st->ref_names.size() - 1; i >= 0; i--)
                if (st->ref_names[i] == chrom) { found = i; break; }
            if (found >= 0) cur = found;
            else {
                const int last = (int)st->ref_names.size() - 1;
                if (prev_end != UINT32_MAX) st->ranges[last].abs_end = prev_end;
                if (want_vars && !vars.vars.empty()) {
                    on_contig(st->ref_names[last], vars, false);
                    vars.clear();
                }
                st->ref_names.push_back(chrom);
                st->ranges.emplace_back();
                cur = (int)st->ref_names.size() - 1;
                prev_end = UINT32_MAX;
            }
        }
        Ranges &R = st->ranges[cur];
        if (fmt == IntervalFormat::VCF) {
            if (want_vars && tok.size() >= 10)
                variant_from_vcf_fields(tok[3], tok[4], tok[8], tok[9], (uint32_t)strtoul(tok[1], nullptr, 10), &vars);
            // insert_vcf_line, blockjoin.c:1348-1430
            if (tok.size() < 2) continue;
            const uint32_t pos = (uint32_t)strtoul(tok[1], nullptr, 10);
            if (prev_end != UINT32_MAX && pos < prev_end) {
                char msg[256];
                snprintf(msg, sizeof(msg), "[E::insert_vcf_line] vcf not sorted? last line pos=%d, current pos=%d", (int)prev_end, (int)pos);
                if (fatal) *fatal = msg;
                return true;
            }
            if (tok.size() < 10) continue;
            const int i_ps = field_index(tok[8], "PS");
            if (i_ps < 0) continue;
            const char *ps;
            int ps_l;
            if (!field_by_index(tok[9], i_ps, &ps, &ps_l)) continue;
            if (ps_l == 1 && ps[0] == '.') continue;
            char num[32];
            snprintf(num, sizeof(num), "%.*s", ps_l > 30 ? 30 : ps_l, ps);
            const uint32_t group = (uint32_t)strtoul(num, nullptr, 10);
            if (prev_group == UINT32_MAX) {
                prev_group = group;
                prev_end = pos;
                R.abs_start = pos;
            }
            if (group == prev_group) prev_end = pos;
            else {
                if (prev_end != UINT32_MAX) {
                    R.starts.push_back(prev_end);
                    R.ends.push_back(group);
                    R.decisions.push_back(-1);
                }
                prev_group = group;
                prev_end = pos;
            }
        } else {
            // insert_gtf_line, blockjoin.c:1305-1345
            const size_t col_s = fmt == IntervalFormat::TSV ? 1 : 3, col_e = fmt == IntervalFormat::TSV ? 2 : 4;
            if (tok.size() > col_s) {
                const uint32_t s = (uint32_t)strtoul(tok[col_s], nullptr, 10);
                if (prev_end != UINT32_MAX) {
                    R.starts.push_back(prev_end);
                    R.ends.push_back(s);
                    R.decisions.push_back(-1);
                } else R.abs_start = s;
            }
            if (tok.size() > col_e) prev_end = (uint32_t)strtoul(tok[col_e], nullptr, 10);
        }
    }
    if (want_vars && cur >= 0) {
        // the reference tags the last contig unconditionally (blockjoin.c:2149-2155)
        on_contig(st->ref_names.back(), vars, true);
        vars.clear();
    }
    if (prev_end != UINT32_MAX && cur >= 0) st->ranges[cur].abs_end = prev_end;
    for (Ranges &R : st->ranges) { R.n = R.starts.size(); R.n_decisions = R.decisions.size(); }
    return true;
}

void store_raw_intervals(Ranges *r) {
    r->raw.clear();
    for (size_t i = 0; i < r->n; i++) r->raw.push_back({r->starts[i], r->ends[i]});
}

void merge_close_intervals(Ranges *r, int threshold) {
    if (r->n <= 1) { r->n_decisions = r->n; return; }
    size_t j = 0;
    for (size_t i = 1; i < r->n; i++) {
        // uint32 arithmetic compared against an int threshold: unsigned (blockjoin.c:2198)
        if ((uint32_t)(r->starts[i] - r->ends[j]) < (uint32_t)threshold) {
            r->dropped.push_back({r->ends[j], r->starts[i]});
            r->ends[j] = r->ends[i];
        } else {
            j++;
            r->starts[j] = r->starts[i];
            r->ends[j] = r->ends[i];
        }
    }
    r->n_decisions = r->n;  // decisions keep the pre-merge count
    r->n = j + 1;           // starts/ends keep their stale tail
}

void lift_decisions(Storage *st) {
    for (Ranges &rr : st->ranges) {
        rr.phaseblocks.clear();
        size_t j = 0;
        for (size_t i = 0; i < rr.n_decisions; i++) {
            if (rr.decisions[i] < 0) {
                while (j < rr.raw.size() && rr.raw[j].e <= rr.ends[i]) {
                    rr.decisions_onraw.push_back(rr.decisions[i]);
                    j++;
                }
            } else {
                if (j < rr.raw.size() && rr.raw[j].e < rr.ends[i]) {
                    size_t j2 = j;
                    for (; j2 < rr.raw.size(); j2++) if (rr.raw[j2].e == rr.ends[i]) break;
                    if (j2 < rr.raw.size()) {  // assert(found) in the reference
                        rr.raw[j].e = rr.ends[i];
                        rr.raw.erase(rr.raw.begin() + (long)j + 1, rr.raw.begin() + (long)j2 + 1);
                    }
                }
                rr.decisions_onraw.push_back(rr.decisions[i]);
                j++;
            }
        }
    }
}

void make_flips_onraw(Storage *st) {
    for (Ranges &rr : st->ranges) {
        int flip = 0;
        for (int d : rr.decisions_onraw) {
            if (d < 0) flip = 0; else flip ^= d;
            rr.flips_onraw.push_back(flip);
        }
    }
}

void generate_new_phase_blocks(Storage *st) {
    for (Ranges &rr : st->ranges) {
        uint32_t start = rr.abs_start, end = UINT32_MAX;
        const size_t N = rr.decisions_onraw.size();
        for (size_t i = 0; i < N; i++) {
            if (rr.decisions_onraw[i] >= 0) continue;
            end = rr.raw[i].s;
            rr.phaseblocks.push_back({start, end});
            start = rr.raw[i].e;
        }
        if (N > 0 && end != rr.abs_end) {
            end = end == UINT32_MAX ? rr.abs_start : end;  // sic: the last block restarts at the gap start
            rr.phaseblocks.push_back({end, rr.abs_end});
        }
    }
}

}  // namespace pomfret
