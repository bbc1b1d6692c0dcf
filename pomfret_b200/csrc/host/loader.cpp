// Host-side read loader, see loader.h.
#include "loader.h"
#include <cstdio>
#include <cstring>

namespace pomfret {

bool BamReader::open(const std::string &path) {
    close();
    fn = path;
    fp = hts_open(path.c_str(), "r");
    if (!fp) return false;
    idx = sam_index_load(fp, path.c_str());
    hdr = sam_hdr_read(fp);
    rec = bam_init1();
    return hdr != nullptr;
}

void BamReader::close() {
    if (rec) bam_destroy1(rec);
    if (idx) hts_idx_destroy(idx);
    if (hdr) sam_hdr_destroy(hdr);
    if (fp) sam_close(fp);
    rec = nullptr; idx = nullptr; hdr = nullptr; fp = nullptr;
}

void WindowReads::clear() {
    descs.clear();
    qname_off.clear();
    qnames.clear();
    arena.clear();
    n_bases = 0;
}

int hp_from_record(const bam1_t *b) {
    uint8_t *t = bam_aux_get(b, "HP");
    if (!t) return kHaptagUnphased;
    int hp = (int)bam_aux2i(t);
    if (hp == 0) {
        fprintf(stderr, "[W::%s] irregular HP tag? qn=%s qs=%d\n", "get_hp_from_aln", bam_get_qname(b), (int)b->core.pos);
        return kHaptagUnphased;
    }
    return hp - 1;
}

void describe_record(const bam1_t *b, int hp, pomfret_gpu_read_desc *d) {
    memset(d, 0, sizeof(*d));
    d->pos = (uint32_t)b->core.pos;
    d->l_qseq = (uint32_t)b->core.l_qseq;
    d->n_cigar = b->core.n_cigar;
    d->flag = b->core.flag;
    d->mapq = b->core.qual;
    d->hp = hp;
    d->mn = -1;
    d->ml_len = -1;
    d->cigar = bam_get_cigar(b);
    d->seq = bam_get_seq(b);
    uint8_t *mm = bam_aux_get(b, "MM");
    if (!mm) mm = bam_aux_get(b, "Mm");
    if (mm) {
        if (mm[0] != 'Z') {  // htslib: "MM tag is not of type Z" => no modifications
            d->tags_malformed = 1;
            d->mm = "";
        } else {
            d->mm = (const char *)mm + 1;
            d->mm_len = (uint32_t)strlen(d->mm);
        }
    }
    uint8_t *mn = bam_aux_get(b, "MN");
    if (mn) {
        int64_t v = bam_aux2i(mn);
        if (v != b->core.l_qseq && b->core.l_qseq) d->tags_malformed = 1;
        d->mn = v >= 0 && v <= INT32_MAX ? (int32_t)v : -1;
    }
    uint8_t *ml = bam_aux_get(b, "ML");
    if (!ml) ml = bam_aux_get(b, "Ml");
    if (ml) {
        if (ml[0] != 'B' || ml[1] != 'C') d->tags_malformed = 1;
        else {
            uint32_t n = (uint32_t)ml[2] | (uint32_t)ml[3] << 8 | (uint32_t)ml[4] << 16 | (uint32_t)ml[5] << 24;
            d->ml = ml + 6;
            d->ml_len = (int32_t)n;
        }
    }
    uint8_t *md = bam_aux_get(b, "MD");
    if (md && md[0] == 'Z') {
        d->md = (const char *)md + 1;
        d->md_len = (uint32_t)strlen(d->md);
    }
}

int for_each_window_record(BamReader &bam, const char *chrom, uint32_t ref_start, uint32_t ref_end, int readlen_threshold,
                           int min_mapq, const RawTagMap *raw_tags, const std::function<void(const bam1_t *, int hp)> &fn) {
    const int itvl_s = (int)ref_start, itvl_e = (int)ref_end;
    // (the reference mallocs strlen(chrom) + 30 here, blockjoin.c:1057: any contig name fits)
    const std::string region = std::string(chrom) + ":" + std::to_string((itvl_s - kReadback) > 0 ? itvl_s - kReadback : 0) + "-" +
                               std::to_string(itvl_e + kReadback);
    hts_itr_t *itr = sam_itr_querys(bam.idx, bam.hdr, region.c_str());
    if (!itr) return POMFRET_GPU_ERR_ARG;
    bam1_t *b = bam.rec;
    while (sam_itr_next(bam.fp, itr, b) >= 0) {
        const int flag = b->core.flag;
        const uint32_t mapq = b->core.qual;
        const uint32_t len = (uint32_t)b->core.l_qseq;
        if ((flag & 4) || (flag & 256) || (flag & 2048)) continue;
        if (mapq < (uint32_t)min_mapq) continue;
        if (len < 2 || len < (uint32_t)readlen_threshold) continue;
        float de = -1;
        uint8_t *t = bam_aux_get(b, "de");
        if (t) de = (float)bam_aux2f(t);
        if (de > kMinAlnDe) continue;
        int hp;
        if (raw_tags) {
            auto it = raw_tags->find(bam_get_qname(b));
            hp = it != raw_tags->end() ? it->second : kHaptagUnphased;
        } else hp = hp_from_record(b);
        fn(b, hp);
    }
    hts_itr_destroy(itr);
    return 0;
}

// Copies only what the engine reads (CIGAR, SEQ, MM, ML, MD): base qualities and the other tags stay behind.
int load_window(BamReader &bam, const char *chrom, uint32_t ref_start, uint32_t ref_end, int readlen_threshold,
                int min_mapq, const RawTagMap *raw_tags, WindowReads *out) {
    out->clear();
    out->ref_start = ref_start;
    out->ref_end = ref_end;
    struct Off { size_t cigar, seq, mm, ml, md; };
    std::vector<Off> offs;
    auto put = [&](const void *p, size_t n) {
        size_t off = (out->arena.size() + 15) & ~(size_t)15;
        out->arena.resize(off + n + 1);
        if (n) memcpy(out->arena.data() + off, p, n);
        out->arena[off + n] = 0;
        return off;
    };
    int rc = for_each_window_record(bam, chrom, ref_start, ref_end, readlen_threshold, min_mapq, raw_tags,
                                    [&](const bam1_t *b, int hp) {
        pomfret_gpu_read_desc d;
        describe_record(b, hp, &d);
        Off o;
        o.cigar = put(d.cigar, (size_t)d.n_cigar * 4);
        o.seq = put(d.seq, ((size_t)d.l_qseq + 1) / 2);
        o.mm = d.mm ? put(d.mm, d.mm_len) : 0;
        o.ml = d.ml_len >= 0 ? put(d.ml, (size_t)d.ml_len) : 0;
        o.md = d.md ? put(d.md, d.md_len) : 0;
        offs.push_back(o);
        out->descs.push_back(d);
        out->qname_off.push_back((uint32_t)out->qnames.size());
        out->qnames.append(bam_get_qname(b));
        out->qnames.push_back('\0');
        out->n_bases += d.l_qseq;
    });
    if (rc) return rc;
    const uint8_t *base = out->arena.data();
    for (size_t i = 0; i < offs.size(); i++) {
        pomfret_gpu_read_desc &d = out->descs[i];
        d.cigar = (const uint32_t *)(base + offs[i].cigar);
        d.seq = base + offs[i].seq;
        if (d.mm) d.mm = (const char *)(base + offs[i].mm);
        if (d.ml_len >= 0) d.ml = base + offs[i].ml;
        if (d.md) d.md = (const char *)(base + offs[i].md);
    }
    return 0;
}

}  // namespace pomfret
