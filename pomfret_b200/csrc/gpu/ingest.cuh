// (a) Compressed ingest: BGZF inflate and BAM record slicing on the device.
//
// The host ships the BGZF blocks of the index chunks of a region query as they lie in the file (a quarter of the
// bytes of the inflated records); it never inflates or parses a record.  Replaces, for that path, what
// sam_itr_next / bam_read1 / bgzf_read do for load_reads_given_interval (reference blockjoin.c:1056-1084) and for
// pre_haplotagging_read_in_one_ref (:1853-1866):
//
//   inflate_kernel      one thread per BGZF block (RFC 1951: stored, fixed and dynamic Huffman blocks; 9-bit /
//                       6-bit first-level tables in local memory, canonical bit-by-bit decode behind them),
//                       ISIZE and CRC-32 (slicing-by-4) checked against the block footer.  Blocks of a stream are
//                       inflated back to back, so records that span blocks are contiguous.
//   walk_kernel         one thread per stream (an index chunk: whole records only): follows the block_size chain,
//                       stops at the first record of another target or at / behind the region end; counts, then
//                       writes the record offsets (two passes around a scan).
//   slice_kernel        one warp per record: core fields, the record filters of blockjoin.c:1081-1084 (flag, MAPQ,
//                       length, `de`), bam_endpos, the aux walk for MM/Mm, ML/Ml, MN, MD, HP, de, CG (long CIGARs,
//                       SAM spec 4.2.2) -> one pomfret_gpu_sliced_record per record with device addresses of its fields.
// The host then decides which windows a record belongs to and hands the records to add_reads_device(): the gather
// kernel copies the fields out of the inflated stream into the aligned batch blob, device to device.
#ifndef POMFRET_GPU_INGEST_CUH
#define POMFRET_GPU_INGEST_CUH
#include "gpu_rt.h"
#include "types.h"
#include "pomfret_gpu.h"

namespace pomfret_gpu {

struct InflateParams {
    const uint8_t *comp;                   // the compressed blocks as they lie in the file
    const pomfret_gpu_bgzf_block *blocks;  // comp_off, csize (whole block), isize, out_off
    uint32_t n_blocks;
    uint8_t *out;
    const uint32_t *crc_tab;               // [4][256] slicing-by-4 tables of the reflected polynomial 0xEDB88320
    int32_t *status;                       // per block: 0 or an error code
    int32_t *n_bad;
    int check_crc;
};

// ---- one warp per BGZF block ----
// Lane 0 walks the Huffman codes (the bit-serial part of DEFLATE) over a window of the compressed bytes that the warp
// keeps staged in shared memory, and turns them into tokens (literal byte / length + distance), 32 at a time; the 32
// lanes then write the literals in one store and copy each match together (a match whose distance is shorter than
// its length repeats a pattern that is already complete in front of it, so every byte of it is independent).
// Code tables live in shared memory too.  The warp finally checks ISIZE and the CRC-32 of what it wrote: 32 segment
// CRCs (slicing-by-4 tables) folded with the x^(8n) mod P operator.
constexpr int INF_LIT_BITS = 9, INF_DIST_BITS = 6;
constexpr int INF_WARPS = 8;             // warps (blocks in flight) per CTA
constexpr uint32_t INF_WIN = 2048;       // staged window of the compressed payload
constexpr uint32_t INF_TOKENS = 32;

struct InflateWarpSmem {
    __align__(16) uint8_t win[INF_WIN + 16];
    uint16_t llut[1 << INF_LIT_BITS], dlut[1 << INF_DIST_BITS], clut[1 << 7];
    uint16_t lcount[16], dcount[16], ccount[16];
    uint16_t lsym[288], dsym[32], csym[19];
    uint8_t lens[320];
    uint32_t tokens[INF_TOKENS];
    uint32_t crc[32];
};

struct BitReader {  // over the staged window; positions are offsets inside the window
    const uint8_t *w;
    uint32_t pos;   // next byte of the window to load
    uint64_t buf;
    int n;
    __device__ __forceinline__ void refill() {
        // at least 32 bits afterwards: four bytes at once (two aligned words of the window, funnel-shifted), then bytes
        // (the window is re-staged long before pos reaches its end)
        if (n <= 32) {
            const uint32_t *w32 = reinterpret_cast<const uint32_t *>(w);
            const uint32_t i = pos >> 2, sh = (pos & 3u) * 8u;
            const uint32_t v = __funnelshift_r(w32[i], w32[i + 1], sh);
            buf |= (uint64_t)v << n;
            n += 32; pos += 4;
        }
        while (n <= 56) { buf |= (uint64_t)w[pos++] << n; n += 8; }
    }
    // enough for one code (15 bits) plus its extra bits (13): cheaper than a full refill between symbols
    __device__ __forceinline__ void need32() {
        if (n < 32) {
            const uint32_t *w32 = reinterpret_cast<const uint32_t *>(w);
            const uint32_t i = pos >> 2, sh = (pos & 3u) * 8u;
            const uint32_t v = __funnelshift_r(w32[i], w32[i + 1], sh);
            buf |= (uint64_t)v << n;
            n += 32; pos += 4;
        }
    }
    __device__ __forceinline__ uint32_t peek(int k) const { return (uint32_t)buf & ((1u << k) - 1u); }
    __device__ __forceinline__ void drop(int k) { buf >>= k; n -= k; }
    __device__ __forceinline__ uint32_t take(int k) { uint32_t v = peek(k); drop(k); return v; }
};

// Canonical Huffman code of `n` symbols with the given lengths: counts per length, symbols in code order, and a
// first-level table of `bits` bits (entry: symbol | length << 12; 0: longer code or unused).  Returns false for an
// over-subscribed set of lengths.  (one lane)
__device__ bool build_code(const uint8_t *lens, int n, uint16_t *count /*[16]*/, uint16_t *symbol, uint16_t *lut, int bits) {
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[lens[i]]++;
    int left = 1;
    for (int l = 1; l < 16; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int i = 0; i < n; i++) if (lens[i]) symbol[offs[lens[i]]++] = (uint16_t)i;
    for (int i = 0; i < (1 << bits); i++) lut[i] = 0;
    // codes in canonical order: length by length, symbols ascending; the table is indexed by the bit-reversed code
    uint32_t code = 0, idx = 0;
    for (int l = 1; l <= bits; l++) {
        for (int k = 0; k < count[l]; k++, code++, idx++) {
            const uint32_t rev = __brev(code) >> (32 - l);
            const uint16_t e = (uint16_t)(symbol[idx] | (l << 12));
            for (uint32_t j = rev; j < (1u << bits); j += 1u << l) lut[j] = e;
        }
        code <<= 1;
    }
    return true;
}

// one symbol: first-level table, else the canonical walk (RFC 1951 3.2.2)
__device__ __forceinline__ int decode_symbol(BitReader &br, const uint16_t *count, const uint16_t *symbol, const uint16_t *lut, int bits) {
    const uint16_t e = lut[br.peek(bits)];
    if (e) { br.drop(e >> 12); return e & 0xfff; }
    int code = 0, first = 0, index = 0;
    uint64_t b = br.buf;
    for (int l = 1; l < 16; l++) {
        code |= (int)(b & 1u);
        b >>= 1;
        const int c = count[l];
        if (code - c < first) { br.drop(l); return symbol[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

// Length / distance bases and extra-bit counts (RFC 1951 3.2.5), symbols 257..285 and 0..29
__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_ext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_ext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// the 32 bits of the staged window that start at bit `bitpos` (two aligned words, funnel-shifted): enough for one code
// (15 bits) with its extra bits (13), so the symbol loop keeps nothing but the bit position between symbols
__device__ __forceinline__ uint32_t win_bits(const uint8_t *win, uint32_t bitpos) {
    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(win);
    const uint32_t i = bitpos >> 5;
    return __funnelshift_r(w32[i], w32[i + 1], bitpos & 31u);
}

// one symbol out of the bits `v`: first-level table, else the canonical walk (RFC 1951 3.2.2); *len = its code length
__device__ __forceinline__ int decode_symbol_bits(uint32_t v, const uint16_t *count, const uint16_t *symbol, const uint16_t *lut, int bits, uint32_t *len) {
    const uint16_t e = lut[v & ((1u << bits) - 1u)];
    if (e) { *len = e >> 12; return e & 0xfff; }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        code |= (int)(v & 1u);
        v >>= 1;
        const int c = count[l];
        if (code - c < first) { *len = (uint32_t)l; return symbol[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    *len = 0;
    return -1;
}

__device__ __forceinline__ uint32_t crc32_update(const uint32_t *tab, uint32_t crc, const uint8_t *p, uint32_t n) {
    crc = ~crc;
    while (n && ((uintptr_t)p & 3u)) { crc = tab[(crc ^ *p++) & 0xffu] ^ (crc >> 8); n--; }
    for (; n >= 4; n -= 4, p += 4) {
        crc ^= *reinterpret_cast<const uint32_t *>(p);
        crc = tab[3 * 256 + (crc & 0xffu)] ^ tab[2 * 256 + ((crc >> 8) & 0xffu)] ^ tab[256 + ((crc >> 16) & 0xffu)] ^ tab[crc >> 24];
    }
    while (n--) crc = tab[(crc ^ *p++) & 0xffu] ^ (crc >> 8);
    return ~crc;
}
// a(x) * b(x) mod P over GF(2), reflected representation (x^0 is bit 31)
__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1u)) == 0) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
// x^(8 n) mod P; x2n[i] = x^(2^i) mod P is tab[1024 + i]
__device__ __forceinline__ uint32_t crc_x8nmodp(const uint32_t *tab, uint32_t n) {
    uint32_t p = 1u << 31, k = 3;
    while (n) { if (n & 1u) p = crc_multmodp(tab[1024 + (k & 31u)], p); n >>= 1; k++; }
    return p;
}

__device__ __forceinline__ uint8_t ld_cg_u8(const uint8_t *p) {
#ifdef POMFRET_CUDA_EMU
    return *p;
#else
    return __ldcg(p);  // bytes this warp wrote a moment ago: read where the stores went (L2)
#endif
}

__global__ void __launch_bounds__(INF_WARPS * 32) inflate_kernel(InflateParams P) {
    __shared__ InflateWarpSmem smem[INF_WARPS];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t bi = blockIdx.x * INF_WARPS + warp;
    if (bi >= P.n_blocks) return;
    InflateWarpSmem &sm = smem[warp];
    const pomfret_gpu_bgzf_block B = P.blocks[bi];
    const uint8_t *blk = P.comp + B.comp_off;
    int err = 0;
    // ---- BGZF header (RFC 1952 member with a BC extra subfield) ----
    uint32_t xlen = 0;
    if (B.csize < 28 || blk[0] != 0x1f || blk[1] != 0x8b || blk[2] != 8 || !(blk[3] & 4)) err = 1;
    else xlen = (uint32_t)blk[10] | ((uint32_t)blk[11] << 8);
    if (!err && 12 + xlen + 8 > B.csize) err = 1;
    uint8_t *out = P.out + B.out_off;
    const uint32_t isize = B.isize;
    uint32_t o = 0;
    if (!err) {
        const uint8_t *pay = blk + 12 + xlen;                 // deflate payload
        const uint32_t pay_len = B.csize - 12 - xlen - 8;
        uint32_t win_off = 0;                                 // payload offset of the window's first byte
        // (re)stage the window at payload offset `at` (all lanes); bytes behind the payload read as zero
        auto stage = [&](uint32_t at) {
            for (uint32_t i = lane; i < INF_WIN + 16; i += 32) sm.win[i] = at + i < pay_len ? pay[at + i] : (uint8_t)0;
            win_off = at;
            __syncwarp();
        };
        stage(0);
        BitReader br;
        br.w = sm.win; br.pos = 0; br.buf = 0; br.n = 0;
        bool last = false;
        while (!last && !err) {
            // ---- block header and code tables (lane 0), the window re-staged first so that 1 KB is ahead ----
            {
                const uint32_t consumed = __shfl_sync(FULL_MASK, br.pos - (uint32_t)(br.n >> 3), 0);  // whole bytes still in the bit buffer are re-read
                const uint32_t bit_rem = __shfl_sync(FULL_MASK, (uint32_t)(br.n & 7), 0);
                if (consumed >= INF_WIN / 2) {
                    // keep the partial byte: restart the reader at the byte that holds the next bit
                    const uint32_t at = win_off + consumed - (bit_rem ? 1u : 0u);
                    stage(at);
                    if (lane == 0) {
                        br.pos = 0; br.buf = 0; br.n = 0;
                        if (bit_rem) { br.refill(); br.drop(8 - (int)bit_rem); }
                    }
                }
            }
            int type = 0;
            uint32_t stored_len = 0, stored_at = 0;
            if (lane == 0) {
                br.refill();
                last = br.take(1) != 0;
                type = (int)br.take(2);
                if (type == 0) {  // stored
                    br.drop(br.n & 7);
                    br.refill();
                    const uint32_t len = br.take(16), nlen = br.take(16);
                    if ((len ^ 0xffffu) != nlen) err = 2;
                    stored_len = len;
                    stored_at = win_off + br.pos - (uint32_t)(br.n >> 3);  // payload offset of the first stored byte
                } else if (type == 3) err = 2;
                else if (type == 1) {  // fixed code
                    for (int i = 0; i < 144; i++) sm.lens[i] = 8;
                    for (int i = 144; i < 256; i++) sm.lens[i] = 9;
                    for (int i = 256; i < 280; i++) sm.lens[i] = 7;
                    for (int i = 280; i < 288; i++) sm.lens[i] = 8;
                    build_code(sm.lens, 288, sm.lcount, sm.lsym, sm.llut, INF_LIT_BITS);
                    for (int i = 0; i < 30; i++) sm.lens[i] = 5;
                    build_code(sm.lens, 30, sm.dcount, sm.dsym, sm.dlut, INF_DIST_BITS);
                } else {  // dynamic code (at most 19*3 + 320*14 bits of header: inside the staged kilobyte)
                    br.refill();
                    const int nlen = (int)br.take(5) + 257, ndist = (int)br.take(5) + 1, ncode = (int)br.take(4) + 4;
                    if (nlen > 286 || ndist > 30) err = 3;
                    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                    for (int i = 0; i < 19; i++) sm.lens[i] = 0;
                    for (int i = 0; i < ncode && !err; i++) { br.refill(); sm.lens[order[i]] = (uint8_t)br.take(3); }
                    if (!err && !build_code(sm.lens, 19, sm.ccount, sm.csym, sm.clut, 7)) err = 3;
                    int idx = 0;
                    while (!err && idx < nlen + ndist) {
                        br.refill();
                        const int sym = decode_symbol(br, sm.ccount, sm.csym, sm.clut, 7);
                        if (sym < 0) { err = 3; break; }
                        if (sym < 16) sm.lens[idx++] = (uint8_t)sym;
                        else {
                            int rep, val = 0;
                            if (sym == 16) { if (idx == 0) { err = 3; break; } val = sm.lens[idx - 1]; rep = 3 + (int)br.take(2); }
                            else if (sym == 17) rep = 3 + (int)br.take(3);
                            else rep = 11 + (int)br.take(7);
                            if (idx + rep > nlen + ndist) { err = 3; break; }
                            while (rep--) sm.lens[idx++] = (uint8_t)val;
                        }
                    }
                    if (!err && sm.lens[256] == 0) err = 3;
                    if (!err) {
                        uint8_t dl[32];
                        for (int i = 0; i < ndist; i++) dl[i] = sm.lens[nlen + i];
                        if (!build_code(sm.lens, nlen, sm.lcount, sm.lsym, sm.llut, INF_LIT_BITS)) err = 3;
                        else if (!build_code(dl, ndist, sm.dcount, sm.dsym, sm.dlut, INF_DIST_BITS)) err = 3;
                    }
                }
            }
            err = __shfl_sync(FULL_MASK, err, 0);
            last = __shfl_sync(FULL_MASK, (int)last, 0) != 0;
            type = __shfl_sync(FULL_MASK, type, 0);
            if (err) break;
            if (type == 0) {
                // ---- stored block: all lanes copy, then the reader restarts behind it ----
                stored_len = __shfl_sync(FULL_MASK, stored_len, 0);
                stored_at = __shfl_sync(FULL_MASK, stored_at, 0);
                if (stored_at + stored_len > pay_len || o + stored_len > isize) { err = 2; break; }
                for (uint32_t i = lane; i < stored_len; i += 32) out[o + i] = pay[stored_at + i];
                o += stored_len;
                stage(stored_at + stored_len);
                if (lane == 0) { br.pos = 0; br.buf = 0; br.n = 0; }
                continue;
            }
            // ---- symbols: bursts of up to 32 tokens by lane 0, executed by the warp ----
            // (inside the loop lane 0 keeps only the bit position in the staged window; the byte-wise reader takes over
            //  again at the end of the block)
            bool end_of_block = false;
            uint32_t bitpos = 0;
            if (lane == 0) bitpos = br.pos * 8u - (uint32_t)br.n;
            while (!end_of_block && !err) {
                {
                    const uint32_t consumed = __shfl_sync(FULL_MASK, bitpos >> 3, 0);
                    if (consumed >= INF_WIN / 2) {  // (a burst consumes at most 32 * 48 bits: the window never runs out)
                        stage(win_off + consumed);
                        bitpos &= 7u;
                    }
                }
                uint32_t n_tok = 0;
                if (lane == 0) {
                    while (n_tok < INF_TOKENS) {
                        uint32_t v = win_bits(sm.win, bitpos), cl;
                        int sym = decode_symbol_bits(v, sm.lcount, sm.lsym, sm.llut, INF_LIT_BITS, &cl);
                        if (sym < 0) { err = 4; break; }
                        if (sym < 256) { sm.tokens[n_tok++] = (uint32_t)sym; bitpos += cl; continue; }
                        if (sym == 256) { end_of_block = true; bitpos += cl; break; }
                        sym -= 257;
                        if (sym >= 29) { err = 4; break; }
                        const uint32_t lext = c_len_ext[sym];
                        const uint32_t len = (uint32_t)c_len_base[sym] + ((v >> cl) & ((1u << lext) - 1u));
                        bitpos += cl + lext;
                        v = win_bits(sm.win, bitpos);
                        const int ds = decode_symbol_bits(v, sm.dcount, sm.dsym, sm.dlut, INF_DIST_BITS, &cl);
                        if (ds < 0 || ds >= 30) { err = 4; break; }
                        const uint32_t dext = c_dist_ext[ds];
                        const uint32_t dist = (uint32_t)c_dist_base[ds] + ((v >> cl) & ((1u << dext) - 1u));
                        bitpos += cl + dext;
                        sm.tokens[n_tok++] = 0x80000000u | (dist << 9) | len;
                    }
                    if (win_off + (bitpos >> 3) > pay_len + 8) err = 4;  // ran past the payload
                }
                __syncwarp();
                n_tok = __shfl_sync(FULL_MASK, n_tok, 0);
                err = __shfl_sync(FULL_MASK, err, 0);
                end_of_block = __shfl_sync(FULL_MASK, (int)end_of_block, 0) != 0;
                if (err) break;
                // ---- execute the tokens: output offsets by a scan, literals at once, matches one after the other ----
                // (measured and rejected, profiles/r02_inflate_variants.txt: producing the burst one output byte per lane —
                //  token search + reference chasing instead of per-match copies — runs at 20 GB/s against 29 GB/s)
                const uint32_t tok = lane < n_tok ? sm.tokens[lane] : 0u;
                const bool is_match = lane < n_tok && (tok >> 31);
                const uint32_t tlen = lane < n_tok ? (is_match ? tok & 0x1ffu : 1u) : 0u;
                const uint32_t incl = warp_inclusive_sum(tlen);
                const uint32_t at = o + incl - tlen;
                const uint32_t total = __shfl_sync(FULL_MASK, incl, 31);
                if (o + total > isize) { err = 5; break; }
                if (lane < n_tok && !is_match) out[at] = (uint8_t)tok;
                unsigned mm = __ballot_sync(FULL_MASK, is_match);
                const uint32_t tdist = (tok >> 9) & 0xffffu;
                if (__any_sync(FULL_MASK, is_match && tdist > at)) { err = 5; break; }
                while (mm) {
                    const int src_lane = __ffs((int)mm) - 1;
                    mm &= mm - 1;
                    __syncwarp();  // what the batch has written so far is visible to the copy
                    const uint32_t dst = __shfl_sync(FULL_MASK, at, src_lane), len = __shfl_sync(FULL_MASK, tlen, src_lane),
                                   dist = __shfl_sync(FULL_MASK, tdist, src_lane);
                    if (dist >= len || dist >= 32u) {
                        // chunks of at most min(32, dist) bytes never read what they write themselves
                        const uint32_t step = dist < 32u ? dist : 32u;
                        for (uint32_t i0 = 0; i0 < len; i0 += step) {
                            const uint32_t i = i0 + lane;
                            uint8_t v = 0;
                            if (lane < step && i < len) v = ld_cg_u8(out + dst + i - dist);
                            if (lane < step && i < len) out[dst + i] = v;
                            if (i0 + step < len && dist < len) __syncwarp();
                        }
                    } else {
                        // a run: the `dist` bytes in front of the match repeat
                        for (uint32_t i = lane; i < len; i += 32) out[dst + i] = ld_cg_u8(out + dst - dist + i % dist);
                    }
                }
                o += total;
                __syncwarp();
            }
            if (lane == 0) {  // back to the byte-wise reader: restart it at the byte that holds the next bit
                br.pos = bitpos >> 3; br.buf = 0; br.n = 0;
                if (bitpos & 7u) { br.refill(); br.drop((int)(bitpos & 7u)); }
            }
        }
        if (!err && o != isize) err = 6;  // ISIZE of the footer
        __syncwarp();
        if (!err && P.check_crc) {
            // 32 segment CRCs, folded left to right: crc(A || B) = crc(A) * x^(8|B|) mod P  xor  crc(B)
            const uint8_t *f = blk + B.csize - 8;
            const uint32_t want = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
            const uint32_t seg = ((isize + 31u) / 32u + 3u) & ~3u;
            const uint32_t a = lane * seg < isize ? lane * seg : isize, z = a + seg < isize ? a + seg : isize;
            sm.crc[lane] = crc32_update(P.crc_tab, 0u, out + a, z - a);
            __syncwarp();
            if (lane == 0) {
                uint32_t crc = 0;
                const uint32_t shift_full = crc_x8nmodp(P.crc_tab, seg);
                for (uint32_t l = 0; l < 32; l++) {
                    const uint32_t la = l * seg < isize ? l * seg : isize, lz = la + seg < isize ? la + seg : isize;
                    if (lz == la) break;
                    const uint32_t sh = lz - la == seg ? shift_full : crc_x8nmodp(P.crc_tab, lz - la);
                    crc = crc_multmodp(sh, crc) ^ sm.crc[l];
                }
                if (crc != want) err = 7;
            }
            err = __shfl_sync(FULL_MASK, err, 0);
        }
    }
    if (lane == 0) {
        P.status[bi] = err;
        if (err) atomicAdd(P.n_bad, 1);
    }
}

// ---- record walk ----
struct WalkParams {
    const pomfret_gpu_bgzf_stream *streams;  // out_off, ubeg, out_bytes (end of the last record), tid, end0
    uint32_t n_streams;
    const uint8_t *out;
    uint32_t *count;        // per stream: records
    const uint32_t *first;  // per stream: index of its first record (exclusive scan of count), fill pass only
    uint64_t *rec_off;      // fill pass: offset of every record in `out`
    uint32_t *rec_stream;
    int32_t *n_bad;
    int fill;
};

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

__global__ void walk_kernel(WalkParams P) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_streams) return;
    const pomfret_gpu_bgzf_stream S = P.streams[s];
    uint64_t off = S.out_off + S.ubeg;
    const uint64_t end = S.out_off + S.out_bytes;
    uint32_t n = 0;
    const uint32_t base = P.fill ? P.first[s] : 0u;
    while (off + 36 <= end) {
        const uint8_t *r = P.out + off;
        const uint32_t block_size = ld_u32(r);
        if (block_size < 32 || off + 4 + block_size > end) { atomicAdd(P.n_bad, 1); break; }  // the chunk does not hold whole records
        const int32_t tid = (int32_t)ld_u32(r + 4), pos = (int32_t)ld_u32(r + 8);
        if (S.tid != POMFRET_GPU_ANY_TID && (tid != S.tid || pos >= (int32_t)S.end0)) break;  // sam_itr_next: the query is over (records are coordinate sorted)
        if (P.fill) { P.rec_off[base + n] = off; P.rec_stream[base + n] = s; }
        n++;
        off += 4 + (uint64_t)block_size;
    }
    if (!P.fill) P.count[s] = n;
}

// exclusive scan of the per-stream counts (one CTA; streams are few thousand at most)
__global__ void scan_counts_kernel(const uint32_t *count, uint32_t *first, uint32_t n, uint32_t *total) {
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_w[33];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? count[i] : 0u;
        const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
        uint32_t incl = warp_inclusive_sum(v);
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = lane < (blockDim.x >> 5) ? s_w[lane] : 0u;
            uint32_t wi = warp_inclusive_sum(w);
            s_w[lane] = wi - w;
            if (lane == 31) s_w[32] = wi;
        }
        __syncthreads();
        if (i < n) first[i] = s_carry + s_w[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

// ---- record slicing ----
struct SliceParams {
    const uint8_t *out;
    const uint64_t *rec_off;
    const uint32_t *rec_stream;
    uint32_t n_records;
    pomfret_gpu_sliced_record *rec;
    pomfret_gpu_ingest_filter flt;
};

constexpr int SLICE_WARPS = 4;

__global__ void __launch_bounds__(SLICE_WARPS * 32) slice_kernel(SliceParams P) {
    const uint32_t ri = blockIdx.x * SLICE_WARPS + (threadIdx.x >> 5);
    if (ri >= P.n_records) return;
    const unsigned lane = lane_id();
    const uint8_t *r = P.out + P.rec_off[ri];
    const uint32_t block_size = ld_u32(r);
    const uint8_t *endp = r + 4 + block_size;
    const uint32_t l_qname = r[12], mapq = r[13];
    uint32_t n_cigar = (uint32_t)r[16] | ((uint32_t)r[17] << 8);
    const uint32_t flag = (uint32_t)r[18] | ((uint32_t)r[19] << 8);
    const uint32_t l_qseq = ld_u32(r + 20);
    const int32_t pos = (int32_t)ld_u32(r + 8);
    const uint8_t *qname = r + 36;
    const uint8_t *cigar = qname + l_qname;
    const uint8_t *seq = cigar + (size_t)n_cigar * 4;
    const uint8_t *aux = seq + ((size_t)l_qseq + 1) / 2 + l_qseq;
    pomfret_gpu_sliced_record R;
    bool bad = aux > endp;  // fields run past the record
    // ---- aux walk (every lane walks the same tags; only the length of a string is found by the lanes together) ----
    const uint8_t *mm = nullptr, *mm_lc = nullptr, *ml = nullptr, *ml_lc = nullptr, *mn = nullptr, *md = nullptr, *hp = nullptr,
                  *de = nullptr, *cg = nullptr;
    uint32_t mm_len = 0, mm_lc_len = 0, md_len = 0;
    for (const uint8_t *p = aux; !bad && p + 3 <= endp;) {
        const uint32_t t0 = p[0], t1 = p[1], ty = p[2];
        const uint8_t *val = p + 2;  // points at the type byte, like bam_aux_get
        const uint8_t *q = p + 3;
        uint32_t zlen = 0;
        switch (ty) {
        case 'A': case 'c': case 'C': q += 1; break;
        case 's': case 'S': q += 2; break;
        case 'i': case 'I': case 'f': q += 4; break;
        case 'd': q += 8; break;
        case 'Z': case 'H': {
            // first NUL at or behind q, 32 bytes per step
            for (;;) {
                const uint8_t *c = q + zlen + lane;
                const unsigned z = __ballot_sync(FULL_MASK, c >= endp || *c == 0);
                if (z) { zlen += (uint32_t)__ffs((int)z) - 1u; break; }
                zlen += 32;
            }
            if (q + zlen >= endp) bad = true;
            q += zlen + 1;
            break;
        }
        case 'B': {
            if (q + 5 > endp) { bad = true; break; }
            const uint32_t sub = q[0], cnt = ld_u32(q + 1);
            const uint32_t es = (sub == 'c' || sub == 'C') ? 1u : (sub == 's' || sub == 'S') ? 2u : (sub == 'i' || sub == 'I' || sub == 'f') ? 4u : 0u;
            if (!es) { bad = true; break; }
            q += 5 + (size_t)cnt * es;
            break;
        }
        default: bad = true; break;
        }
        if (bad || q > endp) { bad = true; break; }
        // (bam_aux_get returns the first tag of a name)
        if (t0 == 'M' && t1 == 'M' && !mm) { mm = val; mm_len = zlen; }
        else if (t0 == 'M' && t1 == 'm' && !mm_lc) { mm_lc = val; mm_lc_len = zlen; }
        else if (t0 == 'M' && t1 == 'L' && !ml) ml = val;
        else if (t0 == 'M' && t1 == 'l' && !ml_lc) ml_lc = val;
        else if (t0 == 'M' && t1 == 'N' && !mn) mn = val;
        else if (t0 == 'M' && t1 == 'D' && !md) { md = val; md_len = zlen; }
        else if (t0 == 'H' && t1 == 'P' && !hp) hp = val;
        else if (t0 == 'd' && t1 == 'e' && !de) de = val;
        else if (t0 == 'C' && t1 == 'G' && !cg) cg = val;
        p = q;
    }
    bool cg_used = false;
    auto aux2i = [](const uint8_t *v, bool *ok) -> int64_t {  // bam_aux2i
        *ok = true;
        switch (v[0]) {
        case 'c': return (int8_t)v[1];
        case 'C': return v[1];
        case 's': return (int16_t)((uint32_t)v[1] | ((uint32_t)v[2] << 8));
        case 'S': return (uint32_t)v[1] | ((uint32_t)v[2] << 8);
        case 'i': return (int32_t)ld_u32(v + 1);
        case 'I': return ld_u32(v + 1);
        default: *ok = false; return 0;
        }
    };
    // ---- long CIGAR in the CG tag (what htslib's bam_read1 restores): placeholder <l_qseq>S<n>N ----
    if (!bad && cg && n_cigar >= 1 && cg[0] == 'B' && (cg[1] == 'I' || cg[1] == 'i')) {
        const uint32_t first = ld_u32(cigar), n_real = ld_u32(cg + 2);
        if ((first & 15u) == 4u && (first >> 4) == l_qseq && n_real >= n_cigar && n_real < (1u << 29) && pos >= 0) {
            cigar = cg + 6;
            n_cigar = n_real;
            cg_used = true;
        }
    }
    // ---- bam_endpos ----
    uint32_t rlen = 0;
    if (!bad)
        for (uint32_t i = lane; i < n_cigar; i += 32) {
            const uint32_t c = ld_u32(cigar + (size_t)i * 4), op = c & 15u;
            if ((0x18du >> op) & 1u) rlen += c >> 4;
        }
    rlen = warp_sum(rlen);
    if (flag & 4u) rlen = 0;
    if (rlen == 0) rlen = 1;
    if (lane != 0) return;
    // ---- record filters (blockjoin.c:1081-1084 / :1862) ----
    float de_v = -1.f;
    if (de) {  // bam_aux2f
        bool ok;
        if (de[0] == 'f') { uint32_t u = ld_u32(de + 1); de_v = __uint_as_float(u); }
        else if (de[0] == 'd') { uint64_t u = (uint64_t)ld_u32(de + 1) | ((uint64_t)ld_u32(de + 5) << 32); de_v = (float)__longlong_as_double((long long)u); }
        else { const int64_t iv = aux2i(de, &ok); de_v = ok ? (float)iv : 0.f; }
    }
    bool keep = !bad && (P.flt.keep_all_flags || !(flag & (4u | 256u | 2048u)));
    if (keep && mapq < P.flt.min_mapq) keep = false;
    if (keep && (l_qseq < P.flt.min_len_floor || l_qseq < P.flt.min_len)) keep = false;
    if (keep && P.flt.check_de && de_v > P.flt.max_de) keep = false;
    // ---- what describe_record() derives ----
    memset(&R, 0, sizeof(R));
    R.pos = (uint32_t)pos; R.end_pos = (uint32_t)pos + rlen; R.l_qseq = l_qseq; R.n_cigar = n_cigar;
    R.flag = (uint16_t)flag; R.mapq = (uint8_t)mapq;
    R.stream = P.rec_stream[ri];
    R.mn = -1; R.ml_len = -1; R.hp = 254; R.hp_irregular = 0;
    R.cigar = (uint64_t)(uintptr_t)cigar; R.seq = (uint64_t)(uintptr_t)seq;
    R.qname_dev = (uint64_t)(uintptr_t)qname; R.l_qname = (uint8_t)l_qname;
    for (uint32_t i = 0; i < sizeof(R.qname) - 1 && i + 1 < l_qname; i++) R.qname[i] = (char)qname[i];
    const uint8_t *m = mm ? mm : mm_lc;
    if (m) {
        if (m[0] != 'Z') { R.tags_malformed = 1; R.mm = (uint64_t)(uintptr_t)(m + 1); R.mm_len = 0; R.has_mm = 1; }
        else { R.mm = (uint64_t)(uintptr_t)(m + 1); R.mm_len = mm ? mm_len : mm_lc_len; R.has_mm = 1; }
    }
    if (mn) {
        bool ok;
        const int64_t v = aux2i(mn, &ok);
        if (v != (int64_t)l_qseq && l_qseq) R.tags_malformed = 1;
        R.mn = v >= 0 && v <= 0x7fffffff ? (int32_t)v : -1;
    }
    const uint8_t *l = ml ? ml : ml_lc;
    if (l) {
        if (l[0] != 'B' || l[1] != 'C') R.tags_malformed = 1;
        else { R.ml = (uint64_t)(uintptr_t)(l + 6); R.ml_len = (int32_t)ld_u32(l + 2); }
    }
    if (md && md[0] == 'Z') { R.md = (uint64_t)(uintptr_t)(md + 1); R.md_len = md_len; }
    if (hp) {  // get_hp_from_aln, blockjoin.c:910-923
        bool ok;
        const int64_t v = aux2i(hp, &ok);
        if (v == 0) R.hp_irregular = 1; else R.hp = (int32_t)(v - 1);
    }
    R.rec_bytes = 4u + block_size;
    R.tid = (int32_t)ld_u32(r + 4);
    if (hp) { R.hp_type = hp[0]; R.hp_off = (uint32_t)(hp - r); }
    R.cg_cigar = cg_used ? 1 : 0;
    R.keep = keep ? 1 : 0;
    R.bad = bad ? 1 : 0;
    P.rec[ri] = R;
}

// ---- coverage estimate (estimate_read_coverage_dirtyfast, blockjoin.c:951-1040) ----
// The reference adds one to bin i/mod for i = start, start+mod, ... < end of every record that passes its filters and
// then only ever uses the SUM of a contig's bins (tot / n_bins), so the bins are not materialised: every kept record
// of the ingest that starts at or behind min_pos (records in front of it belong to the previous slice of the contig)
// contributes the number of its increments that fall on a bin below n_bins.
__global__ void coverage_kernel(const pomfret_gpu_sliced_record *rec, uint32_t n_records, uint32_t min_pos, uint32_t bin_size,
                                uint32_t n_bins, unsigned long long *total) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t inc = 0;
    if (i < n_records && rec[i].keep && rec[i].pos >= min_pos && rec[i].end_pos > rec[i].pos) {
        const uint32_t s = rec[i].pos, e = rec[i].end_pos;
        inc = (e - s + bin_size - 1) / bin_size;                // i = s + j*mod < e
        const uint32_t first_bin = s / bin_size;                // bin of step j: (s + j*mod)/mod = first_bin + j
        const uint32_t room = first_bin < n_bins ? n_bins - first_bin : 0u;
        if (inc > room) inc = room;
    }
    inc = warp_sum(inc);
    if (lane_id() == 0 && inc) atomicAdd(total, (unsigned long long)inc);
}

// ---- output-BAM re-tagging (output_modify_bam, blockjoin.c:3022-3103): one warp per record ----
// The record is copied from the inflated stream to its place in the output stream with the HP tag set like
// bam_aux_update_int(aln, "HP", v) (blockjoin.c:3092) sets it: v < 255 needs one byte (type C), 255 two (type S); an
// integer HP tag that is large enough keeps its size and becomes unsigned, a smaller one grows (the tags behind it move
// up), a record without the tag gets "HP" + type + value appended; block_size follows.
struct RetagParams {
    const uint8_t *in;            // inflated streams
    const uint64_t *rec_off;
    const pomfret_gpu_sliced_record *rec;
    const uint64_t *dst_off;
    const uint8_t *hp_val;
    uint32_t n_records;
    uint8_t *out;
};

// n bytes from src to dst, any alignment, whole warp: 32-bit stores built from two aligned source words
__device__ __forceinline__ void warp_copy_bytes(uint8_t *dst, const uint8_t *src, uint32_t n, unsigned lane) {
    uint32_t head = (4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u;
    if (head > n) head = n;
    if (lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const uint32_t words = n >> 2;
    const uint32_t sh = (uint32_t)((uintptr_t)src & 3u) * 8u;
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>((uintptr_t)src & ~(uintptr_t)3);
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
    for (uint32_t w = lane; w < words; w += 32) {
        const uint32_t lo = s32[w], hi = sh ? s32[w + 1] : 0u;
        d32[w] = __funnelshift_r(lo, hi, sh);
    }
    const uint32_t done = words << 2;
    if (done + lane < n) dst[done + lane] = src[done + lane];
}

__device__ __forceinline__ uint32_t aux_int_size(uint32_t type) {
    return (type == 'c' || type == 'C') ? 1u : (type == 's' || type == 'S') ? 2u : (type == 'i' || type == 'I') ? 4u : 0u;
}

__global__ void __launch_bounds__(SLICE_WARPS * 32) retag_kernel(RetagParams P) {
    const uint32_t ri = blockIdx.x * SLICE_WARPS + (threadIdx.x >> 5);
    if (ri >= P.n_records) return;
    const unsigned lane = lane_id();
    const uint8_t *r = P.in + P.rec_off[ri];
    uint8_t *o = P.out + P.dst_off[ri];
    const uint32_t n = P.rec[ri].rec_bytes, hp_type = P.rec[ri].hp_type, hp_off = P.rec[ri].hp_off;
    const uint32_t val = P.hp_val[ri];
    const uint32_t old_sz = hp_type ? aux_int_size(hp_type) : 0u;
    if (val == 0u || (hp_type && old_sz == 0u)) {  // left as it is (an HP tag that is not an integer: the update fails, the record is written unchanged)
        warp_copy_bytes(o, r, n, lane);
        return;
    }
    uint32_t sz = val < 255u ? 1u : 2u;
    if (!hp_type) {
        warp_copy_bytes(o, r, n, lane);
        __syncwarp();
        if (lane == 0) {
            o[n] = 'H'; o[n + 1] = 'P'; o[n + 2] = sz == 1u ? 'C' : 'S';
            o[n + 3] = (uint8_t)val;
            if (sz == 2u) o[n + 4] = 0;
            const uint32_t bs = n - 4u + 3u + sz;
            o[0] = (uint8_t)bs; o[1] = (uint8_t)(bs >> 8); o[2] = (uint8_t)(bs >> 16); o[3] = (uint8_t)(bs >> 24);
        }
        return;
    }
    if (old_sz >= sz) {
        warp_copy_bytes(o, r, n, lane);
        __syncwarp();
        if (lane == 0) {
            o[hp_off] = old_sz == 1u ? 'C' : (old_sz == 2u ? 'S' : 'I');
            for (uint32_t i = 0; i < old_sz; i++) o[hp_off + 1u + i] = (uint8_t)(val >> (8u * i));
        }
        return;
    }
    // grow: [0, hp_off) | type + value | the tags behind the old value
    const uint32_t grow = sz - old_sz, tail_from = hp_off + 1u + old_sz;
    warp_copy_bytes(o, r, hp_off, lane);
    warp_copy_bytes(o + hp_off + 1u + sz, r + tail_from, n - tail_from, lane);
    __syncwarp();
    if (lane == 0) {
        o[hp_off] = sz == 1u ? 'C' : 'S';
        o[hp_off + 1u] = (uint8_t)val;
        if (sz == 2u) o[hp_off + 2u] = 0;
        const uint32_t bs = n - 4u + grow;
        o[0] = (uint8_t)bs; o[1] = (uint8_t)(bs >> 8); o[2] = (uint8_t)(bs >> 16); o[3] = (uint8_t)(bs >> 24);
    }
}

}  // namespace pomfret_gpu
#endif
