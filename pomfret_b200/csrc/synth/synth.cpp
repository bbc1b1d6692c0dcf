// Synthetic ONT-like data generator for the methphase hot path.
//
// Writes a coordinate-sorted BAM (+BAI) with MM/ML, MD, de, HP/PS tags and a
// whatshap-style phased VCF (+ a truth TSV of the phase-block gaps) so that the
// reference CPU build and the GPU path read identical files.  Everything is a
// pure function of (seed, config): the reference genome, CpG methylation
// classes and variants are position hashes, reads come from a seeded stream.
//
// Model (all knobs in synth_config):
//   genome      random bases, accidental CG suppressed, designated CpG every ~cpg_period bp
//   variants    het SNV / small indel every ~var_period bp, ALT on a hashed haplotype
//   phase sets  lognormal block lengths, gaps between blocks, random orientation per block
//   reads       exponential start spacing for the target depth, lognormal length,
//               substitution / insertion / deletion errors, optional soft clips,
//               CIGAR M/I/D/S only (never H,P,=,X,N: fatal in reference blockjoin.c:776-778)
//   5mC         per-site class {both methylated, both unmethylated, haplotype specific},
//               bimodal ML with flip noise and a mid-range (no-call) fraction
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>
#include <zlib.h>
#include "htslib/sam.h"
#include "synth.h"

namespace {

inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(mix64(seed) | 1) {}
    uint64_t next() {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        return s * 0x2545F4914F6CDD1Dull;
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return n ? (uint32_t)((next() >> 32) * (uint64_t)n >> 32) : 0; }
    bool chance(double p) { return uni() < p; }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    int geometric(double p_continue) {  // >=1
        int n = 1;
        while (n < 64 && chance(p_continue)) n++;
        return n;
    }
};

const char kBases[4] = {'A', 'C', 'G', 'T'};

struct Genome {
    uint64_t seed;
    int cpg_period;
    // reference bases fixed from outside (REF alleles of an imported VCF), keyed by (contig << 40 | pos)
    const std::unordered_map<uint64_t, char> *fixed = nullptr;
    uint64_t key(int contig, int64_t pos, uint64_t salt) const {
        return mix64(seed ^ mix64(((uint64_t)contig << 40) ^ (uint64_t)pos ^ (salt << 56)));
    }
    bool designated(int c, int64_t p) const { return p >= 0 && key(c, p, 1) % (uint64_t)cpg_period == 0; }
    int raw(int c, int64_t p) const { return (int)(key(c, p, 2) & 3); }
    // local function of (p, p-1, p-2): random base with accidental CG removed, designated CpGs forced
    char base(int c, int64_t p) const {
        if (fixed && !fixed->empty()) {
            auto it = fixed->find(((uint64_t)c << 40) | (uint64_t)p);
            if (it != fixed->end()) return it->second;
        }
        if (designated(c, p)) return 'C';
        if (designated(c, p - 1)) return 'G';
        int r = raw(c, p);
        if (r == 2) {  // 'G': would it follow a C?
            bool prev_c = !designated(c, p - 2) && (designated(c, p - 1) || raw(c, p - 1) == 1);
            if (prev_c) return 'A';
        }
        return kBases[r];
    }
    // methylation class of the CpG whose C sits at p: 0 both meth, 1 both unmeth, 2/3 hap0/hap1 specific
    int meth_class(int c, int64_t p, double f_meth, double f_unmeth) const {
        double u = (key(c, p, 3) >> 11) * (1.0 / 9007199254740992.0);
        if (u < f_meth) return 0;
        if (u < f_meth + f_unmeth) return 1;
        return 2 + (int)(key(c, p, 4) & 1);
    }
};

struct Variant {
    int64_t pos;      // 0-based anchor (VCF POS-1)
    int type;         // 0 SNV, 1 DEL, 2 INS
    int len;          // indel length
    int alt_hap;      // true haplotype carrying ALT
    std::string ref, alt;
    int block;        // phase block index or -1 (unphased)
    int kind;         // 0 het phased candidate, 1 hom (1/1), 2 unphased het
};

struct Block {
    int64_t s, e;  // [s,e) span that holds phased variants
    int orient;
};

struct BamRec {
    int32_t tid;
    int32_t pos;
    uint16_t flag;
    uint8_t mapq;
    std::string qname;
    std::vector<uint32_t> cigar;
    std::string seq;  // ACGT text, reference orientation
    std::vector<uint8_t> aux;
};

void aux_put_int(std::vector<uint8_t> &a, const char *tag, int64_t v) {
    a.push_back(tag[0]); a.push_back(tag[1]);
    if (v >= 0 && v < 256) { a.push_back('C'); a.push_back((uint8_t)v); }
    else if (v >= 0 && v < 65536) { a.push_back('S'); a.push_back(v & 0xff); a.push_back((v >> 8) & 0xff); }
    else { a.push_back('i'); for (int i = 0; i < 4; i++) a.push_back((uint8_t)((uint32_t)v >> (8 * i))); }
}
void aux_put_float(std::vector<uint8_t> &a, const char *tag, float f) {
    a.push_back(tag[0]); a.push_back(tag[1]); a.push_back('f');
    uint8_t b[4]; memcpy(b, &f, 4);
    a.insert(a.end(), b, b + 4);
}
void aux_put_str(std::vector<uint8_t> &a, const char *tag, const std::string &s) {
    a.push_back(tag[0]); a.push_back(tag[1]); a.push_back('Z');
    a.insert(a.end(), s.begin(), s.end());
    a.push_back(0);
}
void aux_put_bytes(std::vector<uint8_t> &a, const char *tag, const std::vector<uint8_t> &v) {
    a.push_back(tag[0]); a.push_back(tag[1]); a.push_back('B'); a.push_back('C');
    uint32_t n = (uint32_t)v.size();
    for (int i = 0; i < 4; i++) a.push_back((uint8_t)(n >> (8 * i)));
    a.insert(a.end(), v.begin(), v.end());
}

int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

void to_bam1(const BamRec &r, int qual_mode, Rng &rng, bam1_t *b) {
    size_t l_qname = r.qname.size() + 1;
    size_t extranul = (l_qname % 4) ? 4 - l_qname % 4 : 0;
    size_t l_seq = r.seq.size();
    size_t need = l_qname + extranul + r.cigar.size() * 4 + (l_seq + 1) / 2 + l_seq + r.aux.size();
    if (need > b->m_data) {
        b->data = (uint8_t *)realloc(b->data, need + 64);
        b->m_data = (uint32_t)(need + 64);
    }
    uint8_t *p = b->data;
    memcpy(p, r.qname.c_str(), l_qname);
    p += l_qname;
    for (size_t i = 0; i < extranul; i++) *p++ = 0;
    memcpy(p, r.cigar.data(), r.cigar.size() * 4);
    p += r.cigar.size() * 4;
    for (size_t i = 0; i < l_seq; i += 2) {
        uint8_t hi = seq_nt16_table[(uint8_t)r.seq[i]];
        uint8_t lo = i + 1 < l_seq ? seq_nt16_table[(uint8_t)r.seq[i + 1]] : 0;
        *p++ = (uint8_t)(hi << 4 | lo);
    }
    if (qual_mode == 0) memset(p, 0xff, l_seq);
    else for (size_t i = 0; i < l_seq; i++) p[i] = (uint8_t)(10 + rng.below(30));
    p += l_seq;
    memcpy(p, r.aux.data(), r.aux.size());
    p += r.aux.size();
    b->l_data = (int)(p - b->data);
    b->core.tid = r.tid;
    b->core.pos = r.pos;
    b->core.qual = r.mapq;
    b->core.flag = r.flag;
    b->core.l_qname = (uint16_t)(l_qname + extranul);
    b->core.l_extranul = (uint8_t)extranul;
    b->core.n_cigar = (uint32_t)r.cigar.size();
    b->core.l_qseq = (int32_t)l_seq;
    b->core.mtid = -1;
    b->core.mpos = -1;
    b->core.isize = 0;
    b->core.bin = (uint16_t)reg2bin(r.pos, bam_endpos(b));
}

struct CigarBuilder {
    std::vector<uint32_t> ops;
    void add(int op, uint32_t len) {
        if (!len) return;
        if (!ops.empty() && (int)(ops.back() & 15) == op) ops.back() += len << 4;
        else ops.push_back(len << 4 | (uint32_t)op);
    }
};

struct MdBuilder {
    std::string s;
    int run = 0;
    bool last_del = false;
    void match() { run++; last_del = false; }
    void flush_run() { s += std::to_string(run); run = 0; }
    void mismatch(char refb) { flush_run(); s.push_back(refb); last_del = false; }
    void del(const std::string &refbases) {
        if (last_del && run == 0) s += refbases;
        else { flush_run(); s.push_back('^'); s += refbases; }
        last_del = true;
    }
    std::string finish() { flush_run(); return s; }
};

char other_base(char b, Rng &rng) {
    for (;;) {
        char c = kBases[rng.below(4)];
        if (c != b) return c;
    }
}

}  // namespace

extern "C" void pomfret_synth_default_config(synth_config *c) {
    memset(c, 0, sizeof(*c));
    c->seed = 20;
    c->coverage = 30.0;
    c->read_len_mean = 20000.0;
    c->read_len_sigma = 0.35;
    c->read_len_min = 2000;
    c->cpg_period = 105;
    c->var_period = 1500;
    c->indel_var_frac = 0.1;
    c->block_len_median = 500000.0;
    c->block_len_sigma = 0.8;
    c->short_block_frac = 0.1;
    c->gap_min = 20000;
    c->gap_max = 150000;
    c->err_rate = 0.01;
    c->softclip_frac = 0.1;
    c->frac_meth = 0.70;
    c->frac_unmeth = 0.15;
    c->ml_flip = 0.08;
    c->ml_mid = 0.06;
    c->hp_drop = 0.10;
    c->tagged = 1;
    c->qual_mode = 0;
    c->compress_level = 1;
    c->frac_low_mapq = 0.02;
    c->frac_high_de = 0.01;
    c->frac_secondary = 0.01;
    c->frac_no_mm = 0.02;
    c->frac_two_segments = 0.2;
    c->frac_multicode = 0.05;
    c->frac_noncpg_calls = 0.0;
    c->n_header_contigs_before = 0;
    c->frac_cpg_listed = 1.0;
    c->de_cap = 0.0;
    c->vcf_in = nullptr;
}

namespace {

struct ContigPlan {
    std::string name;
    int64_t len;
    int tid;
    int64_t region_beg, region_end;  // reads are simulated only inside this span
};

void build_blocks(const synth_config &cfg, const ContigPlan &ct, Rng &rng, std::vector<Block> &blocks) {
    int64_t p = ct.region_beg + 1000 + rng.below(20000);
    while (p < ct.region_end - 50000) {
        double len;
        if (rng.chance(cfg.short_block_frac)) len = 5000 + rng.below(35000);
        else len = cfg.block_len_median * std::exp(cfg.block_len_sigma * rng.normal());
        if (len < 3000) len = 3000;
        Block b;
        b.s = p;
        b.e = std::min<int64_t>(p + (int64_t)len, ct.region_end - 1000);
        b.orient = (int)(rng.next() & 1);
        if (b.e - b.s < 3000) break;
        blocks.push_back(b);
        int64_t gap = cfg.gap_min + rng.below((uint32_t)(cfg.gap_max - cfg.gap_min + 1));
        p = b.e + gap;
    }
}

void build_variants(const synth_config &cfg, const Genome &g, const ContigPlan &ct, const std::vector<Block> &blocks,
                    Rng &rng, std::vector<Variant> &vars) {
    size_t bi = 0;
    int64_t p = ct.region_beg + 50 + rng.below((uint32_t)cfg.var_period);
    while (p < ct.region_end - 100) {
        Variant v;
        v.pos = p;
        v.block = -1;
        while (bi < blocks.size() && blocks[bi].e <= p) bi++;
        if (bi < blocks.size() && blocks[bi].s <= p && p < blocks[bi].e) v.block = (int)bi;
        v.alt_hap = (int)(rng.next() & 1);
        double u = rng.uni();
        v.kind = 0;
        if (u < 0.05) v.kind = 1;       // homozygous ALT
        else if (u < 0.12) v.kind = 2;  // het left unphased by the caller
        if (v.block < 0 && v.kind == 0) v.kind = 2;
        bool indel = rng.chance(cfg.indel_var_frac);
        char refb = g.base(ct.tid, p);
        if (!indel) {
            v.type = 0; v.len = 1;
            v.ref = std::string(1, refb);
            v.alt = std::string(1, other_base(refb, rng));
        } else if (rng.chance(0.5)) {
            v.type = 1; v.len = 1 + (int)rng.below(3);
            v.ref = std::string(1, refb);
            for (int k = 1; k <= v.len; k++) v.ref.push_back(g.base(ct.tid, p + k));
            v.alt = std::string(1, refb);
        } else {
            v.type = 2; v.len = 1 + (int)rng.below(3);
            v.ref = std::string(1, refb);
            v.alt = v.ref;
            for (int k = 0; k < v.len; k++) v.alt.push_back(kBases[rng.below(4)]);
        }
        vars.push_back(v);
        // keep variants apart so that alleles never overlap
        p += 8 + (int64_t)(-std::log(1.0 - rng.uni() * 0.999999) * cfg.var_period);
    }
    // every block needs at least two phased variants to define its span; shrink blocks to their variants
}

// ---- imported VCF (--vcf-in): header contigs, variants and phase sets of a real call set ----
struct VcfRecord { int64_t pos; std::string ref, alt, gt; int64_t ps; };
struct VcfImport {
    std::vector<std::string> contig_names;
    std::vector<uint32_t> contig_lens;
    std::unordered_map<std::string, std::vector<VcfRecord>> recs;
};

bool read_vcf(const char *fn, VcfImport *out) {
    gzFile f = gzopen(fn, "rb");
    if (!f) return false;
    std::string line;
    char buf[1 << 16];
    auto next_line = [&]() -> bool {
        line.clear();
        while (gzgets(f, buf, sizeof(buf))) {
            line += buf;
            if (!line.empty() && line.back() == '\n') { line.pop_back(); return true; }
        }
        return !line.empty();
    };
    while (next_line()) {
        if (line.rfind("##contig=<", 0) == 0) {
            size_t a = line.find("ID="), b = line.find("length=");
            if (a == std::string::npos || b == std::string::npos) continue;
            size_t ae = line.find_first_of(",>", a);
            out->contig_names.push_back(line.substr(a + 3, ae - a - 3));
            out->contig_lens.push_back((uint32_t)strtoul(line.c_str() + b + 7, nullptr, 10));
            continue;
        }
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> col;
        size_t p = 0;
        while (col.size() < 10) {
            size_t q = line.find('\t', p);
            col.push_back(line.substr(p, q == std::string::npos ? std::string::npos : q - p));
            if (q == std::string::npos) break;
            p = q + 1;
        }
        if (col.size() < 10) continue;
        VcfRecord r;
        r.pos = atoll(col[1].c_str()) - 1;
        r.ref = col[3]; r.alt = col[4];
        r.ps = -1;
        // FORMAT keys -> sample values
        std::vector<std::string> keys, vals;
        for (int which = 0; which < 2; which++) {
            const std::string &src = col[8 + which];
            std::vector<std::string> &dst = which ? vals : keys;
            size_t a = 0;
            for (;;) {
                size_t b = src.find(':', a);
                dst.push_back(src.substr(a, b == std::string::npos ? std::string::npos : b - a));
                if (b == std::string::npos) break;
                a = b + 1;
            }
        }
        for (size_t i = 0; i < keys.size() && i < vals.size(); i++) {
            if (keys[i] == "GT") r.gt = vals[i];
            if (keys[i] == "PS" && vals[i] != ".") r.ps = atoll(vals[i].c_str());
        }
        out->recs[col[0]].push_back(r);
    }
    gzclose(f);
    return true;
}

// blocks = phase sets (span of their phased records), variants = the simple records (SNV, anchored indels)
void import_variants(const std::vector<VcfRecord> &recs, Rng &rng, std::vector<Block> &blocks, std::vector<Variant> &vars) {
    std::vector<int64_t> ps_ids;
    auto phased = [](const VcfRecord &r) { return r.ps >= 0 && r.gt.size() == 3 && r.gt[1] == '|' && r.gt[0] != r.gt[2]; };
    for (const VcfRecord &r : recs) {
        if (!phased(r)) continue;
        size_t k = std::find(ps_ids.begin(), ps_ids.end(), r.ps) - ps_ids.begin();
        // (whatshap names a phase set after the position of its first variant: the set starts there even if the
        //  file at hand is a slice that begins later)
        if (k == ps_ids.size()) { ps_ids.push_back(r.ps); blocks.push_back(Block{std::min(r.pos, r.ps - 1), r.pos + 1, (int)(rng.next() & 1)}); }
        blocks[k].s = std::min(blocks[k].s, r.pos);
        blocks[k].e = std::max(blocks[k].e, r.pos + 1);
    }
    std::vector<size_t> order(blocks.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return blocks[a].s < blocks[b].s; });
    std::vector<Block> sorted;
    std::vector<int> remap(blocks.size(), -1);
    for (size_t i : order) {
        if (!sorted.empty() && blocks[i].s < sorted.back().e) continue;  // interleaved phase sets: keep the first
        remap[i] = (int)sorted.size();
        sorted.push_back(blocks[i]);
    }
    int64_t prev_end = -100;
    for (const VcfRecord &r : recs) {
        if (r.alt.find(',') != std::string::npos || r.ref.empty() || r.alt.empty()) continue;
        bool ok = true;
        for (char c : r.ref + r.alt) if (c != 'A' && c != 'C' && c != 'G' && c != 'T') ok = false;
        if (!ok) continue;
        Variant v;
        v.pos = r.pos;
        if (r.ref.size() == 1 && r.alt.size() == 1) { v.type = 0; v.len = 1; }
        else if (r.alt.size() == 1 && r.ref[0] == r.alt[0]) { v.type = 1; v.len = (int)r.ref.size() - 1; }
        else if (r.ref.size() == 1 && r.ref[0] == r.alt[0]) { v.type = 2; v.len = (int)r.alt.size() - 1; }
        else continue;
        if (v.pos <= prev_end + 8) continue;  // alleles must not overlap
        v.ref = r.ref; v.alt = r.alt;
        v.block = -1;
        v.alt_hap = (int)(rng.next() & 1);
        const bool hom = r.gt.size() == 3 && r.gt[0] == '1' && r.gt[2] == '1';
        const bool het = r.gt.size() == 3 && ((r.gt[0] == '0' && r.gt[2] == '1') || (r.gt[0] == '1' && r.gt[2] == '0'));
        if (hom) v.kind = 1;
        else if (het && phased(r)) {
            size_t k = std::find(ps_ids.begin(), ps_ids.end(), r.ps) - ps_ids.begin();
            if (remap[k] < 0) v.kind = 2;
            else { v.kind = 0; v.block = remap[k]; v.alt_hap = (r.gt[0] == '1' ? 0 : 1) ^ sorted[(size_t)remap[k]].orient; }
        } else if (het) v.kind = 2;
        else continue;
        vars.push_back(v);
        prev_end = v.pos + (v.type == 1 ? v.len : 0);
    }
    blocks.swap(sorted);
}

struct ReadSim {
    const synth_config &cfg;
    const Genome &g;
    const ContigPlan &ct;
    const std::vector<Variant> &vars;
    const std::vector<Block> &blocks;

    // simulate one read; returns false if it cannot be placed
    bool make(Rng &rng, int64_t start, int64_t ref_len, uint64_t serial, BamRec &out) const {
        if (start + ref_len + 8 >= ct.len) ref_len = ct.len - start - 8;
        if (ref_len < 200) return false;
        int hap = (int)(rng.next() & 1);
        bool rev = (rng.next() & 1) != 0;
        std::string &seq = out.seq;
        seq.clear();
        std::vector<int64_t> seq2ref;
        seq2ref.reserve((size_t)ref_len + 512);
        CigarBuilder cg;
        MdBuilder md;
        int n_err = 0;

        int clipL = 0, clipR = 0;
        if (rng.chance(cfg.softclip_frac)) clipL = 1 + (int)rng.below(200);
        if (rng.chance(cfg.softclip_frac)) clipR = 1 + (int)rng.below(200);
        for (int i = 0; i < clipL; i++) { seq.push_back(kBases[rng.below(4)]); seq2ref.push_back(-1); }
        cg.add(BAM_CSOFT_CLIP, (uint32_t)clipL);

        size_t vi = std::lower_bound(vars.begin(), vars.end(), start,
                                     [](const Variant &v, int64_t p) { return v.pos < p; }) - vars.begin();
        const int64_t end = start + ref_len;
        int64_t r = start;
        while (r < end) {
            bool edge = (r - start) < 12 || (end - r) < 12;
            char refb = g.base(ct.tid, r);
            // germline variant anchored here?
            while (vi < vars.size() && vars[vi].pos < r) vi++;
            if (!edge && vi < vars.size() && vars[vi].pos == r) {
                const Variant &v = vars[vi];
                bool carries = v.kind == 1 || v.alt_hap == hap;
                if (carries && r + v.len + 12 < end) {
                    if (v.type == 0) {
                        seq.push_back(v.alt[0]); seq2ref.push_back(r);
                        cg.add(BAM_CMATCH, 1); md.mismatch(refb);
                        r++;
                    } else if (v.type == 1) {
                        seq.push_back(refb); seq2ref.push_back(r);
                        cg.add(BAM_CMATCH, 1); md.match();
                        cg.add(BAM_CDEL, (uint32_t)v.len);
                        md.del(v.ref.substr(1));
                        r += 1 + v.len;
                    } else {
                        seq.push_back(refb); seq2ref.push_back(r);
                        cg.add(BAM_CMATCH, 1); md.match();
                        for (int k = 1; k <= v.len; k++) { seq.push_back(v.alt[k]); seq2ref.push_back(-1); }
                        cg.add(BAM_CINS, (uint32_t)v.len);
                        r++;
                    }
                    continue;
                }
            }
            if (!edge && rng.chance(cfg.err_rate)) {
                double u = rng.uni();
                n_err++;
                if (u < 0.4) {
                    seq.push_back(other_base(refb, rng)); seq2ref.push_back(r);
                    cg.add(BAM_CMATCH, 1); md.mismatch(refb);
                    r++;
                } else if (u < 0.7) {
                    int l = rng.geometric(0.3);
                    for (int k = 0; k < l; k++) { seq.push_back(kBases[rng.below(4)]); seq2ref.push_back(-1); }
                    cg.add(BAM_CINS, (uint32_t)l);
                    n_err += l - 1;
                } else {
                    int l = rng.geometric(0.3);
                    if (r + l + 12 >= end) l = 1;
                    std::string d;
                    for (int k = 0; k < l; k++) d.push_back(g.base(ct.tid, r + k));
                    cg.add(BAM_CDEL, (uint32_t)l); md.del(d);
                    r += l;
                    n_err += l - 1;
                }
                continue;
            }
            seq.push_back(refb); seq2ref.push_back(r);
            cg.add(BAM_CMATCH, 1); md.match();
            r++;
        }
        for (int i = 0; i < clipR; i++) { seq.push_back(kBases[rng.below(4)]); seq2ref.push_back(-1); }
        cg.add(BAM_CSOFT_CLIP, (uint32_t)clipR);

        out.tid = ct.tid;
        out.pos = (int32_t)start;
        out.flag = rev ? BAM_FREVERSE : 0;
        out.mapq = rng.chance(cfg.frac_low_mapq) ? (uint8_t)rng.below(10) : 60;
        if (rng.chance(cfg.frac_secondary)) out.flag |= (rng.next() & 1) ? BAM_FSECONDARY : BAM_FSUPPLEMENTARY;
        out.cigar = cg.ops;
        char name[64];
        uint64_t h1 = mix64(cfg.seed ^ (serial * 0x9E37ull + 77)), h2 = mix64(h1);
        snprintf(name, sizeof(name), "%08x-%04x-%04x-%04x-%012llx", (uint32_t)(h1 >> 32), (uint32_t)(h1 >> 16) & 0xffff,
                 (uint32_t)h1 & 0xffff, (uint32_t)(h2 >> 48), (unsigned long long)(h2 & 0xffffffffffffull));
        out.qname = name;

        // ---- aux tags ----
        out.aux.clear();
        int64_t aligned = (int64_t)seq.size() - clipL - clipR;
        float de = aligned > 0 ? (float)n_err / (float)aligned : 0.f;
        if (cfg.de_cap > 0 && de > (float)cfg.de_cap) de = (float)cfg.de_cap;
        if (rng.chance(cfg.frac_high_de)) de = 0.11f + 0.2f * (float)rng.uni();
        aux_put_int(out.aux, "NM", n_err);
        aux_put_float(out.aux, "de", de);
        aux_put_str(out.aux, "MD", md.finish());

        // haplotag from the overlapped phase block
        if (cfg.tagged) {
            int best = -1;
            int64_t best_ov = 0;
            size_t b0 = std::lower_bound(blocks.begin(), blocks.end(), start,
                                         [](const Block &b, int64_t p) { return b.e <= p; }) - blocks.begin();
            for (size_t b = b0; b < blocks.size() && blocks[b].s < end; b++) {
                int64_t ov = std::min(end, blocks[b].e) - std::max(start, blocks[b].s);
                if (ov > best_ov) { best_ov = ov; best = (int)b; }
            }
            if (best >= 0 && best_ov >= 2000 && !rng.chance(cfg.hp_drop)) {
                aux_put_int(out.aux, "HP", (hap ^ blocks[best].orient) + 1);
                aux_put_int(out.aux, "PS", blocks[best].s + 1);
            }
        }

        // ---- MM / ML ----
        if (!rng.chance(cfg.frac_no_mm)) {
            const int n = (int)seq.size();
            std::vector<uint32_t> deltas;
            std::vector<uint8_t> ml_m, ml_h;
            uint32_t skipped = 0;
            auto emit = [&](int64_t cpos, bool is_cpg_context) {
                // cpos: reference position of the CpG's C (or -1)
                int q;
                if (cpos >= 0 && is_cpg_context && g.designated(ct.tid, cpos)) {
                    int cls = g.meth_class(ct.tid, cpos, cfg.frac_meth, cfg.frac_unmeth);
                    bool meth = cls == 0 ? true : cls == 1 ? false : (cls - 2) == hap;
                    if (rng.chance(cfg.ml_flip)) meth = !meth;
                    if (rng.chance(cfg.ml_mid)) q = 100 + (int)rng.below(56);
                    else q = meth ? 200 + (int)rng.below(56) : (int)rng.below(41);
                } else q = (int)rng.below(256);
                deltas.push_back(skipped);
                skipped = 0;
                ml_m.push_back((uint8_t)q);
                ml_h.push_back((uint8_t)rng.below(q < 128 ? 40 : 256 - q));
            };
            if (!rev) {
                for (int i = 0; i < n; i++) {
                    if (seq[i] != 'C') continue;
                    bool cpg = i + 1 < n && seq[i + 1] == 'G';
                    if ((cpg && (cfg.frac_cpg_listed >= 1.0 || rng.chance(cfg.frac_cpg_listed))) || (!cpg && rng.chance(cfg.frac_noncpg_calls))) emit(seq2ref[i], cpg);
                    else skipped++;
                }
            } else {
                for (int i = n - 1; i >= 0; i--) {
                    if (seq[i] != 'G') continue;
                    bool cpg = i > 0 && seq[i - 1] == 'C';
                    if ((cpg && (cfg.frac_cpg_listed >= 1.0 || rng.chance(cfg.frac_cpg_listed))) || (!cpg && rng.chance(cfg.frac_noncpg_calls))) emit(seq2ref[i] >= 0 ? seq2ref[i] - 1 : -1, cpg);
                    else skipped++;
                }
            }
            std::string dl;
            dl.reserve(deltas.size() * 3);
            for (uint32_t d : deltas) { dl.push_back(','); dl += std::to_string(d); }
            double u = rng.uni();
            std::string mm;
            std::vector<uint8_t> ml;
            if (u < cfg.frac_multicode) {  // C+hm? : shared positions, interleaved ML (h first)
                mm = "C+hm?" + dl + ";";
                for (size_t i = 0; i < ml_m.size(); i++) { ml.push_back(ml_h[i]); ml.push_back(ml_m[i]); }
            } else if (u < cfg.frac_multicode + cfg.frac_two_segments) {
                mm = "C+h?" + dl + ";C+m?" + dl + ";";
                ml = ml_h;
                ml.insert(ml.end(), ml_m.begin(), ml_m.end());
            } else {
                mm = std::string("C+m") + (rng.chance(0.9) ? "?" : ".") + dl + ";";
                ml = ml_m;
            }
            aux_put_str(out.aux, "MM", mm);
            aux_put_bytes(out.aux, "ML", ml);
            if (rng.chance(0.5)) aux_put_int(out.aux, "MN", n);
        }
        return true;
    }
};

}  // namespace

extern "C" int pomfret_synth_write(const synth_config *cfgp, const char *const *contig_names,
                                   const int64_t *contig_lens, const int64_t *region_beg, const int64_t *region_end,
                                   int n_contigs, const char *prefix) {
    const synth_config &cfg = *cfgp;
    std::unordered_map<uint64_t, char> fixed_bases;
    Genome g{cfg.seed, cfg.cpg_period, &fixed_bases};
    VcfImport imported;
    if (cfg.vcf_in && !read_vcf(cfg.vcf_in, &imported)) return -7;
    std::string fn_bam = std::string(prefix) + ".bam";
    std::string fn_vcf = std::string(prefix) + ".vcf.gz";
    std::string fn_truth = std::string(prefix) + ".truth.tsv";

    // header: optional filler contigs first so that tids are not trivially 0
    std::vector<ContigPlan> plan;
    sam_hdr_t hdr;
    memset(&hdr, 0, sizeof(hdr));
    int n_fill = cfg.n_header_contigs_before;
    hdr.n_targets = n_fill + n_contigs;
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    if (cfg.vcf_in) {  // the header is the call set's contig list; -C picks contigs of it
        n_fill = 0;
        names = imported.contig_names;
        lens = imported.contig_lens;
        hdr.n_targets = (int)names.size();
    }
    for (int i = 0; i < n_fill; i++) { names.push_back("fill" + std::to_string(i + 1)); lens.push_back(1000000); }
    for (int i = 0; i < n_contigs; i++) {
        ContigPlan p;
        p.name = contig_names[i];
        p.len = contig_lens[i];
        p.tid = n_fill + i;
        if (cfg.vcf_in) {
            size_t k = std::find(names.begin(), names.end(), p.name) - names.begin();
            if (k == names.size()) return -8;
            p.tid = (int)k;
            p.len = lens[k];
        } else {
            names.push_back(contig_names[i]);
            lens.push_back((uint32_t)contig_lens[i]);
        }
        p.region_beg = region_beg ? std::max<int64_t>(0, region_beg[i]) : 0;
        p.region_end = region_end && region_end[i] > 0 ? std::min<int64_t>(region_end[i], p.len) : p.len;
        plan.push_back(p);
    }
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    std::vector<char *> name_ptrs;
    for (size_t i = 0; i < names.size(); i++) {
        text += "@SQ\tSN:" + names[i] + "\tLN:" + std::to_string(lens[i]) + "\n";
        name_ptrs.push_back(const_cast<char *>(names[i].c_str()));
    }
    text += "@PG\tID:pomfret-synth\tPN:pomfret-synth\n";
    hdr.l_text = text.size();
    hdr.text = const_cast<char *>(text.c_str());
    hdr.target_name = name_ptrs.data();
    hdr.target_len = lens.data();

    char mode[8];
    snprintf(mode, sizeof(mode), "w%d", cfg.compress_level);
    BGZF *out = bgzf_open(fn_bam.c_str(), mode);
    if (!out) return -1;
    if (bam_hdr_write(out, &hdr) != 0) return -2;

    gzFile vcf = gzopen(cfg.vcf_in ? "/dev/null" : fn_vcf.c_str(), "wb1");
    FILE *truth = fopen(fn_truth.c_str(), "w");
    if (!vcf || !truth) return -3;
    gzprintf(vcf, "##fileformat=VCFv4.2\n##source=pomfret-synth\n");
    for (size_t i = 0; i < names.size(); i++) gzprintf(vcf, "##contig=<ID=%s,length=%u>\n", names[i].c_str(), lens[i]);
    gzprintf(vcf, "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n");
    gzprintf(vcf, "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype quality\">\n");
    gzprintf(vcf, "##FORMAT=<ID=PS,Number=1,Type=Integer,Description=\"Phase set\">\n");
    gzprintf(vcf, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n");
    fprintf(truth, "#chrom\tgap_start\tgap_end\ttruth\n");

    // Planning pass (sequential, cheap): blocks, variants, VCF and truth lines of every contig, and how many read
    // serial numbers each contig consumes (read names and per-read streams derive from one running counter).
    // The reads themselves are then simulated by one thread per contig into BGZF part files that are
    // concatenated in contig order: the records are exactly those a sequential run would write.
    struct ContigJob {
        std::vector<Block> blocks;
        std::vector<Variant> vars;
        Rng rng{0};           // stream state where the read loop starts
        uint64_t serial0 = 0; // first serial number of the contig
        std::string part;     // BGZF part file
        uint64_t n_written = 0, n_bases = 0;
        int rc = 0;
        struct IdxRec { int64_t beg, end; uint64_t off0, off1; };  // offsets inside the part file
        std::vector<IdxRec> idx;
    };
    std::vector<ContigJob> jobs(plan.size());
    uint64_t serial = 0;
    for (size_t ci = 0; ci < plan.size(); ci++) {
        const ContigPlan &ct = plan[ci];
        ContigJob &J = jobs[ci];
        Rng rng(cfg.seed * 1000003ull + ci * 7919ull + 13);
        std::vector<Block> &blocks = J.blocks;
        std::vector<Variant> &vars = J.vars;
        if (cfg.vcf_in) {
            import_variants(imported.recs[ct.name], rng, blocks, vars);
            for (const Variant &v : vars)
                for (size_t k = 0; k < v.ref.size(); k++) fixed_bases[((uint64_t)ct.tid << 40) | (uint64_t)(v.pos + (int64_t)k)] = v.ref[k];
        } else {
            build_blocks(cfg, ct, rng, blocks);
            build_variants(cfg, g, ct, blocks, rng, vars);
        }

        // VCF + truth
        {
            std::vector<int64_t> ps(blocks.size(), -1), last(blocks.size(), -1);
            for (const Variant &v : vars)
                if (v.kind == 0 && v.block >= 0) {
                    if (ps[v.block] < 0) ps[v.block] = v.pos + 1;
                    last[v.block] = v.pos + 1;
                }
            int prev = -1;
            for (size_t bi = 0; bi < blocks.size(); bi++) {
                if (ps[bi] < 0) continue;
                if (prev >= 0)
                    fprintf(truth, "%s\t%lld\t%lld\t%s\n", ct.name.c_str(), (long long)last[prev], (long long)ps[bi],
                            blocks[prev].orient == blocks[bi].orient ? "cis" : "trans");
                prev = (int)bi;
            }
            for (const Variant &v : vars) {
                if (v.kind == 1)
                    gzprintf(vcf, "%s\t%lld\t.\t%s\t%s\t30\tPASS\t.\tGT:GQ\t1/1:30\n", ct.name.c_str(),
                             (long long)v.pos + 1, v.ref.c_str(), v.alt.c_str());
                else if (v.kind == 2 || v.block < 0)
                    gzprintf(vcf, "%s\t%lld\t.\t%s\t%s\t30\tPASS\t.\tGT:GQ\t0/1:30\n", ct.name.c_str(),
                             (long long)v.pos + 1, v.ref.c_str(), v.alt.c_str());
                else {
                    int alt_col = v.alt_hap ^ blocks[v.block].orient;  // VCF haplotype column carrying ALT
                    gzprintf(vcf, "%s\t%lld\t.\t%s\t%s\t30\tPASS\t.\tGT:GQ:PS\t%s:30:%lld\n", ct.name.c_str(),
                             (long long)v.pos + 1, v.ref.c_str(), v.alt.c_str(), alt_col == 0 ? "1|0" : "0|1",
                             (long long)ps[v.block]);
                }
            }
        }
        J.rng = rng;
        J.serial0 = serial;
        J.part = fn_bam + ".part" + std::to_string(ci);
        // dry run of the read loop's own draws (start spacing, length): how many serial numbers it takes
        const double mean_gap = cfg.read_len_mean / cfg.coverage;
        const double mu = std::log(cfg.read_len_mean) - 0.5 * cfg.read_len_sigma * cfg.read_len_sigma;
        double pos = (double)ct.region_beg - cfg.read_len_mean;
        for (;;) {
            pos += -std::log(1.0 - rng.uni() * 0.999999999) * mean_gap;
            if (pos >= (double)ct.region_end) break;
            (void)std::exp(mu + cfg.read_len_sigma * rng.normal());
            serial++;
        }
    }
    if (bgzf_close(out) != 0) return -5;  // header part (ends with the BGZF end-of-file marker)
    gzclose(vcf);
    fclose(truth);

    auto simulate = [&](size_t ci) {
        const ContigPlan &ct = plan[ci];
        ContigJob &J = jobs[ci];
        BGZF *po = bgzf_open(J.part.c_str(), mode);
        if (!po) { J.rc = -1; return; }
        bam1_t *b = bam_init1();
        ReadSim sim{cfg, g, ct, J.vars, J.blocks};
        Rng rng = J.rng;
        uint64_t ser = J.serial0;
        double mean_gap = cfg.read_len_mean / cfg.coverage;
        double mu = std::log(cfg.read_len_mean) - 0.5 * cfg.read_len_sigma * cfg.read_len_sigma;
        double pos = (double)ct.region_beg - cfg.read_len_mean;  // lead-in so that depth is flat at region_beg
        BamRec rec;
        for (;;) {
            pos += -std::log(1.0 - rng.uni() * 0.999999999) * mean_gap;
            if (pos >= (double)ct.region_end) break;
            int64_t len = (int64_t)std::exp(mu + cfg.read_len_sigma * rng.normal());
            if (len < cfg.read_len_min) len = cfg.read_len_min;
            ser++;
            if (pos < 0 || pos < (double)ct.region_beg - 3 * cfg.read_len_mean) continue;
            int64_t start = (int64_t)pos;
            if (start < 0) continue;
            Rng rr(cfg.seed ^ mix64(ser * 0x51ull + ci));
            if (!sim.make(rr, start, len, ser, rec)) continue;
            to_bam1(rec, cfg.qual_mode, rr, b);
            // (what bam_write1 does first: a record starts a new block if it does not fit the current one)
            if (bgzf_flush_try(po, 4 + 32 + (ssize_t)b->l_data - b->core.l_extranul + (b->core.n_cigar > 0xffffu ? 16 : 0)) != 0) { J.rc = -4; break; }
            const uint64_t off0 = (uint64_t)bgzf_tell(po);
            if (bam_write1(po, b) < 0) { J.rc = -4; break; }
            J.idx.push_back({b->core.pos, bam_endpos(b), off0, (uint64_t)bgzf_tell(po)});
            J.n_written++;
            J.n_bases += rec.seq.size();
        }
        bam_destroy1(b);
        if (bgzf_close(po) != 0 && !J.rc) J.rc = -5;
    };
    {
        unsigned hw = std::thread::hardware_concurrency();
        size_t n_thr = std::max<size_t>(1, std::min<size_t>(plan.size(), hw ? hw : 1));
        if (const char *e = getenv("POMFRET_SYNTH_THREADS")) n_thr = std::max<size_t>(1, std::min<size_t>(plan.size(), (size_t)atoi(e)));
        std::atomic<size_t> next(0);
        auto body = [&]() { for (size_t ci; (ci = next.fetch_add(1)) < plan.size();) simulate(ci); };
        std::vector<std::thread> th;
        for (size_t t = 1; t < n_thr; t++) th.emplace_back(body);
        body();
        for (auto &t : th) t.join();
    }
    // concatenate: header blocks, then every part, each without its end-of-file marker; one marker at the very end
    uint64_t n_written = 0, n_bases = 0;
    pomfret_bai_builder *bai = nullptr;
    {
        const size_t kEof = 28;
        FILE *fo = fopen(fn_bam.c_str(), "r+b");
        if (!fo) return -5;
        bai = pomfret_bai_new(hdr.n_targets);
        fseeko(fo, 0, SEEK_END);
        off_t hdr_end = ftello(fo) - (off_t)kEof;
        uint8_t eof_block[28];
        fseeko(fo, hdr_end, SEEK_SET);
        if (fread(eof_block, 1, kEof, fo) != kEof) { fclose(fo); return -5; }
        fseeko(fo, hdr_end, SEEK_SET);
        std::vector<uint8_t> buf((size_t)8 << 20);
        for (ContigJob &J : jobs) {
            if (J.rc) { fclose(fo); return J.rc; }
            FILE *fi = fopen(J.part.c_str(), "rb");
            if (!fi) { fclose(fo); return -5; }
            fseeko(fi, 0, SEEK_END);
            off_t left = ftello(fi) - (off_t)kEof;
            fseeko(fi, 0, SEEK_SET);
            const uint64_t shift = (uint64_t)ftello(fo) << 16;  // virtual offset = compressed offset << 16 | offset in block
            // a reader names the position behind a record that ends its block as the start of the next block:
            // that is the next record's own start (for the last record: where the part's end-of-file marker sat)
            for (size_t i = 0; i < J.idx.size(); i++)
                J.idx[i].off1 = i + 1 < J.idx.size() ? J.idx[i + 1].off0 : (uint64_t)left << 16;
            const int tid = plan[(size_t)(&J - jobs.data())].tid;
            for (const ContigJob::IdxRec &r : J.idx)
                if (pomfret_bai_add(bai, tid, r.beg, r.end, 0, r.off0 + shift, r.off1 + shift) != 0) { fclose(fi); fclose(fo); return -6; }

            while (left > 0) {
                size_t n = fread(buf.data(), 1, (size_t)std::min<off_t>(left, (off_t)buf.size()), fi);
                if (n == 0 || fwrite(buf.data(), 1, n, fo) != n) { fclose(fi); fclose(fo); return -5; }
                left -= (off_t)n;
            }
            fclose(fi);
            remove(J.part.c_str());
            n_written += J.n_written;
            n_bases += J.n_bases;
        }
        if (fwrite(eof_block, 1, kEof, fo) != kEof) { fclose(fo); return -5; }
        off_t total = ftello(fo);
        fclose(fo);
        if (truncate(fn_bam.c_str(), total) != 0) return -5;
    }
    std::string fn_bai = fn_bam + ".bai";
    int rc = getenv("POMFRET_SYNTH_REINDEX") ? (pomfret_bai_finish(bai, nullptr), sam_index_build3(fn_bam.c_str(), fn_bai.c_str(), 0, 1))
                                             : pomfret_bai_finish(bai, fn_bai.c_str());
    if (rc != 0) return -6;
    fprintf(stderr, "[pomfret-synth] wrote %llu reads, %llu bases to %s\n", (unsigned long long)n_written,
            (unsigned long long)n_bases, fn_bam.c_str());
    return 0;
}
