// Host half of the compressed ingest, see ingest.h.
#include "ingest.h"
#include <unistd.h>
#include <cstring>

namespace pomfret {

bool ingest_plan_region(hts_idx_t *idx, int tid, int64_t beg0, int64_t end0, uint32_t run, uint64_t file_size, IngestPlan *plan) {
    hts_itr_t *itr = sam_itr_queryi(idx, tid, beg0, end0);
    if (!itr) return false;
    const uint64_t *pairs = nullptr;
    const int n = pomfret_itr_chunks(itr, &pairs);
    for (int i = 0; i < n; i++) {
        const uint64_t vbeg = pairs[2 * i], vend = pairs[2 * i + 1];
        if (vend <= vbeg) continue;
        IngestPlan::Range r;
        r.file_off = vbeg >> 16;
        // every block that starts in front of the chunk's end, and the end's own block if the chunk ends inside it:
        // its size is not known yet, so the range takes the largest block there can be
        uint64_t stop = (vend >> 16) + ((vend & 0xffff) ? 65536 : 0);
        if (stop > file_size) stop = file_size;
        if (stop <= r.file_off) continue;
        r.bytes = stop - r.file_off;
        r.comp_off = (plan->comp_bytes + 15) & ~(size_t)15;
        r.vbeg = vbeg; r.vend = vend; r.run = run; r.tid = tid; r.end0 = (uint32_t)end0;
        plan->comp_bytes = r.comp_off + r.bytes;
        plan->ranges.push_back(r);
    }
    hts_itr_destroy(itr);
    return true;
}

bool ingest_read(int fd, IngestPlan *plan, uint8_t *comp, std::string *err) {
    uint64_t out_off = 0;
    for (const IngestPlan::Range &r : plan->ranges) {
        uint64_t got = 0;
        while (got < r.bytes) {
            const ssize_t k = pread(fd, comp + r.comp_off + got, r.bytes - got, (off_t)(r.file_off + got));
            if (k <= 0) { if (err) *err = "short read of the BAM file"; return false; }
            got += (uint64_t)k;
        }
        // walk the BGZF headers of the range (RFC 1952 member, extra subfield BC = BSIZE)
        const uint64_t coff_end = r.vend >> 16, uend = r.vend & 0xffff;
        pomfret_gpu_bgzf_stream S;
        memset(&S, 0, sizeof(S));
        S.out_off = (out_off + 15) & ~(uint64_t)15;
        out_off = S.out_off;
        S.ubeg = (uint32_t)(r.vbeg & 0xffff);
        S.tid = r.tid; S.end0 = r.end0;
        S.first_block = (uint32_t)plan->blocks.size();
        uint64_t p = 0, stream_bytes = 0;
        bool closed = false;
        while (p + 18 <= r.bytes) {
            const uint64_t coff = r.file_off + p;
            if (coff > coff_end || (coff == coff_end && uend == 0)) { closed = true; break; }
            const uint8_t *h = comp + r.comp_off + p;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { if (err) *err = "not a BGZF block where the index points"; return false; }
            const uint32_t xlen = (uint32_t)h[10] | ((uint32_t)h[11] << 8);
            int bsize = -1;
            for (uint32_t o = 0; o + 4 <= xlen && 12 + o + 4 <= r.bytes - p;) {
                const uint32_t slen = (uint32_t)h[12 + o + 2] | ((uint32_t)h[12 + o + 3] << 8);
                if (h[12 + o] == 'B' && h[12 + o + 1] == 'C' && slen == 2) bsize = (int)((uint32_t)h[12 + o + 4] | ((uint32_t)h[12 + o + 5] << 8));
                o += 4 + slen;
            }
            if (bsize < 0 || p + (uint64_t)bsize + 1 > r.bytes) { if (err) *err = "truncated BGZF block"; return false; }
            const uint32_t csize = (uint32_t)bsize + 1;
            const uint8_t *f = h + csize - 4;
            const uint32_t isize = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
            if (isize > 65536) { if (err) *err = "BGZF block larger than 64 KiB"; return false; }
            pomfret_gpu_bgzf_block B;
            B.comp_off = r.comp_off + p; B.csize = csize; B.isize = isize; B.out_off = out_off;
            plan->blocks.push_back(B);
            if (coff == coff_end) { stream_bytes += uend; out_off += isize; closed = true; break; }  // the chunk ends inside this block
            stream_bytes += isize;
            out_off += isize;
            p += csize;
        }
        if (!closed && (r.file_off + p < coff_end || uend != 0)) { if (err) *err = "index chunk runs past the blocks read"; return false; }
        S.n_blocks = (uint32_t)plan->blocks.size() - S.first_block;
        S.out_bytes = stream_bytes;
        if (S.ubeg > S.out_bytes) { if (err) *err = "index chunk starts behind its end"; return false; }
        plan->streams.push_back(S);
        plan->stream_run.push_back(r.run);
    }
    return true;
}

}  // namespace pomfret
