// (b) Per-read decode kernel: MM/ML base-modification tags + CIGAR -> reference-coordinate 5mC
// calls at CpG sites.  One warp per alignment record at a time (persistent warps over a queue, longest first).
//
// Replaces, for records that passed the host filters:
//   bam_parse_basemod / bam_mods_at_next_pos as used at reference blockjoin.c:807,832-882
//     (htslib semantics restated from the SAMtags spec),
//   fill_read_meth_record_from_bam_line  blockjoin.c:794-908  (C+m only, CpG check on SEQ, ML -> category),
//   get_mod_poss_on_ref                  blockjoin.c:605-792  (CIGAR walk with its quirks, App. A.3),
//   bam_endpos                           (blockjoin.c:1125).
//
// Data flow per warp (all global accesses are 128-bit or lane-consecutive):
//   1. CIGAR pre-scan: reference span, first stopping op, fatal-op test.
//   2. MM structure scan: ';' positions -> segment table; headers parsed by lane 0.
//   3. Delta lists: 16 characters per lane, classified with SWAR byte masks; every comma's number is
//      converted without a loop from a shared-memory staged chunk; warp prefix sums turn deltas into
//      ranks ("index among the canonical bases").
//   4. SEQ, pass 1: two 512 B tiles per warp step (next step prefetched): nibble match flags, popcounts,
//      one packed warp scan; the rank of every 32-base chunk's first canonical base goes to shared memory.
//      Pass 2, dense over the listed bases (32 per pass): binary search of the chunk holding the rank,
//      select-nth inside the chunk, CpG context on SEQ, ML byte -> category, coalesced stores.
//      Reverse-strand records scan SEQ from its end so ranks count from the read's own 5' end.
//   5. CIGAR walk: M/I ops compacted (read-end, offset) into shared memory in chunks; each kept mod
//      binary-searches the op that the reference's inclusive trigger loop would handle it under;
//      positions de-duplicated with the reference's overwrite rule and written out.
// Records with implicit canonical calls (a listed cytosine outside a CpG, blockjoin.c:666-700) stay on these paths: the
// walk over CIGAR operations and kept mods is then run sequentially by the warp with the lanes scanning the stretches
// for CpGs (implicit_walk).  Records that need the general sequential semantics (several C+m segments, >10 mod streams,
// more segments than the table holds) are collected and run one per thread (decode_generic_kernel).
#ifndef POMFRET_GPU_DECODE_CUH
#define POMFRET_GPU_DECODE_CUH
#include "gpu_rt.h"
#include "types.h"

namespace pomfret_gpu {

constexpr int DEC_WARPS = 4;
constexpr int DEC_MAXSEG = 12;
constexpr int DEC_MI_CAP = 256;       // M/I ops staged per chunk
constexpr int DEC_MM_CHUNK = 512;     // MM bytes staged per step
constexpr int DEC_T = 2;              // SEQ tiles (512 B = 1024 bases each) per scan step
constexpr int DEC_FC = 768;           // chunks (32 bases each) per SEQ section: 24 tiles, 24576 bases
constexpr int N_MODS_LIMIT = 10;      // reference N_MODS, blockjoin.c:34

struct DecodeParams {
    const ReadRec *reads;
    uint32_t n_reads;
    uint32_t n_queue;      // entries of `order` to decode (records shared between windows are decoded once)
    const uint8_t *blob;
    uint32_t *calls_pos;
    uint8_t *calls_cat;
    uint32_t *tmp_rank;   // scratch, same slot layout as the call arrays
    uint32_t *tmp_mpos;
    uint8_t *tmp_mcat;
    uint32_t *r_ncalls, *r_status, *r_end;
    uint32_t *n_overflow;  // records that ran out of call slots (the engine re-runs them with more room)
    uint32_t *next;        // work queue head (zeroed before the launch)
    const uint32_t *order; // queue order: record indices, longest first (nullptr: batch order)
    uint32_t lo, hi;
    uint32_t no_lean;      // test / measurement hook: every record takes the streaming path
    uint32_t *generic_list;  // records that need the general sequential path, handed to decode_generic_kernel (nullptr: lane 0 runs them in place)
    uint32_t *n_generic;
};

struct SegInfo {
    uint32_t list_begin;  // offset of the first ',' (or of ';' for an empty list)
    uint32_t list_end;    // offset of the terminating ';'
    uint32_t n_delta;
    uint32_t total;       // sum of (delta+1)
    uint32_t ml_base;
    uint16_t n_codes;
    int16_t m_idx;        // index of code 'm' or -1
    uint8_t canon;        // 4-bit code of the canonical base (A1 C2 G4 T8 N15)
    uint8_t m_count;      // how many times 'm' occurs in the code list
    uint8_t pad[2];
};

// Whole-record tables of the lean path (records of at most 65535 bases and LEAN_MI M/I operations: every position
// fits 16 bits, so both tables of a record stay in shared memory and nothing goes through scratch arrays).
// (measured, profiles/r02e_decode_variants.txt: 9 / 10 / 12 CTAs per SM with 56 / 48 / 40 registers and smaller tables are
//  slower than 8 CTAs with 64 registers — 0.590 / 0.649 / 0.845 against 0.576 ms on 50 k records)
#ifndef POMFRET_DEC_LEAN_CH
#define POMFRET_DEC_LEAN_CH 2048
#endif
#ifndef POMFRET_DEC_LEAN_MI
#define POMFRET_DEC_LEAN_MI 512
#endif
#ifndef POMFRET_DEC_MIN_CTAS
#define POMFRET_DEC_MIN_CTAS 8
#endif
constexpr int LEAN_CH = POMFRET_DEC_LEAN_CH;   // 16-byte SEQ chunks of a record (a multiple of 128): 2048 = 65536 bases
constexpr int LEAN_MI = POMFRET_DEC_LEAN_MI;   // M/I operations in front of the first stopping operation (a multiple of 8)
constexpr int LEAN_MAXLEN = LEAN_CH * 32 - 1 < 65535 ? LEAN_CH * 32 - 1 : 65535;
static_assert(LEAN_CH % 128 == 0 && LEAN_CH <= 2048 && LEAN_MI % 8 == 0, "lean tables: whole super-tiles, 16-bit positions");
constexpr int16_t LEAN_DROP = (int16_t)0x8000;

struct DecodeWarpSmem {
    __align__(16) uint8_t mmbuf[DEC_MM_CHUNK + 16];
    union {
        struct {                          // lean path
            uint16_t l_first[LEAN_CH + 8];    // rank of the first canonical base of every chunk in scan order; [n_ch] = total
            uint16_t l_end[LEAN_MI];          // read offset behind every M/I operation
            int16_t l_off[LEAN_MI];           // reference offset of the bases of an M operation, LEAN_DROP for I
        };
        union {                           // streaming path (any record size)
            uint32_t first[DEC_FC + 1];   // SEQ scan: rank of the first canonical base of every 32-base chunk of the section
            struct {                      // CIGAR walk (afterwards): staged M/I ops
                uint32_t mi_end[DEC_MI_CAP];
                int32_t mi_off[DEC_MI_CAP];
            };
        };
    };
    SegInfo seg[DEC_MAXSEG];
};

constexpr int32_t MI_DROP = (int32_t)0x80000000;
// Deltas and their running sums saturate here: far beyond any SEQ length the engine accepts (add_read
// rejects l_qseq >= 2^28), so a saturated rank is simply "never found" — what the reference's 64-bit
// arithmetic yields for such lists.
constexpr uint32_t DEC_SAT = 0x40000000u;
__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) {  // a, b <= DEC_SAT
    uint32_t s = a + b;
    return s > DEC_SAT ? DEC_SAT : s;
}
__device__ __forceinline__ uint32_t warp_inclusive_sat_sum(uint32_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL_MASK, v, o);
        if (lane_id() >= (unsigned)o) v = sat_add(v, t);
    }
    return v;
}

__device__ __forceinline__ uint32_t seq_nib(const uint8_t *seq, uint32_t i) {
    return (seq[i >> 1] >> ((~i & 1u) << 2)) & 0xfu;
}

// flags word: bit 4*i set iff nibble i of the 32-bit SEQ word equals the nibble replicated in `pat`
// (nibble i holds base i^1: BAM packs the first base of a byte high).  Zero padding never matches.
__device__ __forceinline__ uint32_t nib_eq_flags(uint32_t w, uint32_t pat) {
    uint32_t x = w ^ pat;
    x |= x >> 1;
    x |= x >> 2;
    return ~x & 0x11111111u;
}
// prmt.b32 with its hardware selector semantics: nibble i of `sel` (low 16 bits) picks byte (nibble & 7) of {hi:lo};
// bit 3 of the nibble replicates that byte's sign bit instead.
__device__ __forceinline__ uint32_t prmt_raw(uint32_t lo, uint32_t hi, uint32_t sel) {
#ifdef POMFRET_CUDA_EMU
    const uint64_t src = ((uint64_t)hi << 32) | lo;
    uint32_t out = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t n = (sel >> (4 * i)) & 15u;
        uint32_t byte = (uint32_t)(src >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        out |= byte << (8 * i);
    }
    return out;
#else
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
#endif
}
// Number of nibbles of a 16-byte SEQ chunk equal to the base code C (2) or G (4), without a popcount: the SEQ
// nibbles themselves are used as byte selectors into a table that holds 1 at the wanted code's index (codes with
// bit 3 replicate the sign of a table byte, which is 0), four nibbles per instruction; the 0/1 bytes of the eight
// results add up in byte lanes.
__device__ __forceinline__ uint32_t count_code_in_chunk(const uint4 &v, uint32_t tab_lo, uint32_t tab_hi) {
    uint32_t acc = prmt_raw(tab_lo, tab_hi, v.x) + prmt_raw(tab_lo, tab_hi, v.x >> 16);
    acc += prmt_raw(tab_lo, tab_hi, v.y) + prmt_raw(tab_lo, tab_hi, v.y >> 16);
    acc += prmt_raw(tab_lo, tab_hi, v.z) + prmt_raw(tab_lo, tab_hi, v.z >> 16);
    acc += prmt_raw(tab_lo, tab_hi, v.w) + prmt_raw(tab_lo, tab_hi, v.w >> 16);
    return (acc * 0x01010101u) >> 24;
}

// base index (0..7) of the n-th (0-based, base order) flagged base of a flags word
__device__ __forceinline__ uint32_t select_base_in_word(uint32_t f, uint32_t n) {
    uint32_t g = ((f & 0x01010101u) << 4) | ((f >> 4) & 0x01010101u);  // bit 4b <-> base b
    uint32_t b = 0;
    uint32_t c = (uint32_t)__popc(g & 0xffffu);
    if (n >= c) { n -= c; g >>= 16; b = 4; }
    c = (uint32_t)__popc(g & 0xffu);
    if (n >= c) { n -= c; g >>= 8; b += 2; }
    c = (uint32_t)__popc(g & 0xfu);
    if (n >= c) b += 1;
    return b;
}
// number of SEQ nibbles equal to `code` (code != 0), whole warp
__device__ uint32_t count_base(const uint8_t *seq, uint32_t n_bytes, uint32_t code) {
    const uint32_t pat = code * 0x11111111u;
    uint32_t c = 0;
    for (uint32_t off = lane_id() * 16; off < n_bytes; off += 512) {
        uint4 v = *reinterpret_cast<const uint4 *>(seq + off);
        c += __popc(nib_eq_flags(v.x, pat)) + __popc(nib_eq_flags(v.y, pat)) + __popc(nib_eq_flags(v.z, pat)) +
             __popc(nib_eq_flags(v.w, pat));
    }
    return warp_sum(c);
}

__device__ __forceinline__ int base_code_of(int ch) {
    switch (ch) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': case 'U': return 8; case 'N': return 15;
    default: return -1;
    }
}
__device__ __forceinline__ bool is_digit(int c) { return c >= '0' && c <= '9'; }
__device__ __forceinline__ bool is_alpha(int c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
__device__ __forceinline__ uint32_t comp_code(uint32_t c) {
    // complement of a 4-bit base code: reverse the 4 bits (A1<->T8, C2<->G4, N15 fixed)
    return ((c & 1u) << 3) | ((c & 2u) << 1) | ((c & 4u) >> 1) | ((c & 8u) >> 3);
}

// SWAR byte classification of four characters at a time
__device__ __forceinline__ uint32_t byte_eq_mask4(uint32_t w, uint32_t pat) {  // 0x80 in every byte of w equal to pat's
    const uint32_t x = w ^ pat;
    const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | x | 0x7f7f7f7fu);
}
__device__ __forceinline__ uint32_t nondigit_mask4(uint32_t w) {  // 0x80 in every byte outside '0'..'9'
    const uint32_t y = w ^ 0x30303030u;
    return (((y & 0x7f7f7f7fu) + 0x76767676u) | y) & 0x80808080u;
}
__device__ __forceinline__ uint32_t pack4(uint32_t m80) {  // 0x80 flags of four bytes -> four bits
    return (((m80 >> 7) * 0x00204081u) >> 21) & 0xfu;
}

// ---------------------------------------------------------------------------------------------
// MM structure scan.  Returns (uniform across the warp) the number of segments, or -1 malformed,
// or -2 "take the generic path".  Segment table in sm.seg.
// ---------------------------------------------------------------------------------------------
__device__ int mm_scan_segments(const uint8_t *mm, uint32_t mm_len, DecodeWarpSmem &sm) {
    const unsigned lane = lane_id();
    if (mm_len == 0) return 0;
    // pass 1: positions of ';'
    int n_seg = 0;
    bool too_many = false;
    for (uint32_t base = 0; base < mm_len; base += 32 * 16) {
        uint32_t off = base + lane * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (off < mm_len) v = *reinterpret_cast<const uint4 *>(mm + off);  // blob segments are padded to 16 B
        uint32_t semi = pack4(byte_eq_mask4(v.x, 0x3b3b3b3bu)) | pack4(byte_eq_mask4(v.y, 0x3b3b3b3bu)) << 4 |
                        pack4(byte_eq_mask4(v.z, 0x3b3b3b3bu)) << 8 | pack4(byte_eq_mask4(v.w, 0x3b3b3b3bu)) << 12;
        semi &= off < mm_len ? (mm_len - off >= 16u ? 0xffffu : (1u << (mm_len - off)) - 1u) : 0u;  // characters of the string only
        uint32_t cnt = __popc(semi);
        uint32_t incl = warp_inclusive_sum(cnt);
        uint32_t excl = incl - cnt;
        uint32_t tot = __shfl_sync(FULL_MASK, incl, 31);
        uint32_t idx = n_seg + excl;
        while (semi) {
            int b = __ffs(semi) - 1;
            semi &= semi - 1;
            if (idx < DEC_MAXSEG) sm.seg[idx].list_end = off + b;
            idx++;
        }
        n_seg += (int)tot;
        if (n_seg > DEC_MAXSEG) too_many = true;
    }
    __syncwarp();
    if (too_many) return -2;
    // the string must end with ';' (htslib: "Missing semicolon")
    if (n_seg == 0 || sm.seg[n_seg - 1].list_end != mm_len - 1) return -1;
    // pass 2: headers, lane 0
    int err = 0;
    if (lane == 0) {
        uint32_t p = 0;
        int n_streams = 0;
        for (int s = 0; s < n_seg && !err; s++) {
            SegInfo &g = sm.seg[s];
            const uint32_t e = g.list_end;
            int canon = base_code_of(mm[p]);
            if (canon < 0) { err = -1; break; }
            p++;
            if (p > e || (mm[p] != '+' && mm[p] != '-')) { err = -1; break; }
            p++;
            int n_codes = 0, m_idx = -1, m_count = 0;
            if (p <= e && is_digit(mm[p])) {
                while (p <= e && is_digit(mm[p])) p++;
                n_codes = 1;
            } else {
                while (p <= e && is_alpha(mm[p])) {
                    if (mm[p] == 'm') { if (m_idx < 0) m_idx = n_codes; m_count++; }
                    n_codes++;
                    p++;
                }
            }
            if (p <= e && (mm[p] == '.' || mm[p] == '?')) p++;
            else if (p > e || (mm[p] != ',' && mm[p] != ';')) { err = -1; break; }
            if (p > e || (mm[p] != ',' && mm[p] != ';')) { err = -1; break; }
            if (mm[p] == ';' && p != e) { err = -1; break; }  // cannot happen: e is the first ';' after the start
            if (n_codes > 0 && n_streams + n_codes >= 256) { err = -1; break; }
            n_streams += n_codes;
            g.list_begin = p;
            g.canon = (uint8_t)canon;
            g.n_codes = (uint16_t)n_codes;
            g.m_idx = (int16_t)m_idx;
            g.m_count = (uint8_t)m_count;
            g.n_delta = 0;
            g.total = 0;
            p = e + 1;
        }
        if (!err && n_streams > N_MODS_LIMIT) err = -2;  // a position could carry more than N_MODS entries
    }
    err = __shfl_sync(FULL_MASK, err, 0);
    __syncwarp();
    if (err) return err;
    return n_seg;
}

// ---------------------------------------------------------------------------------------------
// Parse the delta list of one segment: count the deltas, sum (delta+1) and, if rank_out != nullptr,
// write the cumulative target index c_k = sum_{j<=k}(delta_j+1) - 1 of every delta.
// Returns false (uniform) if the list is malformed.
// ---------------------------------------------------------------------------------------------
__device__ bool mm_parse_list(const uint8_t *mm, SegInfo &g, DecodeWarpSmem &sm, uint32_t *rank_out, uint32_t rank_cap) {
    const unsigned lane = lane_id();
    const uint32_t lb = g.list_begin, le = g.list_end;  // list chars are [lb, le), le is ';'
    uint32_t n_delta = 0, total = 0;
    bool bad = false;
    for (uint32_t base = lb & ~15u; base < le; base += DEC_MM_CHUNK) {
        // stage [base, base+512+16) — bytes past the MM string are blob padding / later fields, never used
        uint32_t off = base + lane * 16;
        uint4 v = *reinterpret_cast<const uint4 *>(mm + off);
        *reinterpret_cast<uint4 *>(sm.mmbuf + lane * 16) = v;
        if (lane == 0) *reinterpret_cast<uint4 *>(sm.mmbuf + DEC_MM_CHUNK) = *reinterpret_cast<const uint4 *>(mm + base + DEC_MM_CHUNK);
        __syncwarp();
        // byte classes of this lane's 16 characters as bit masks (SWAR, no per-character loop)
        const uint32_t comma16 = pack4(byte_eq_mask4(v.x, 0x2c2c2c2cu)) | pack4(byte_eq_mask4(v.y, 0x2c2c2c2cu)) << 4 |
                                 pack4(byte_eq_mask4(v.z, 0x2c2c2c2cu)) << 8 | pack4(byte_eq_mask4(v.w, 0x2c2c2c2cu)) << 12;
        const uint32_t nd16 = pack4(nondigit_mask4(v.x)) | pack4(nondigit_mask4(v.y)) << 4 | pack4(nondigit_mask4(v.z)) << 8 |
                              pack4(nondigit_mask4(v.w)) << 12;
        const uint32_t lo_i = lb > off ? (lb - off < 16u ? lb - off : 16u) : 0u;
        const uint32_t hi_i = le > off ? (le - off < 16u ? le - off : 16u) : 0u;
        const uint32_t below_hi = (1u << hi_i) - 1u;                          // positions before the terminating ';'
        const uint32_t range = hi_i > lo_i ? below_hi & ~((1u << lo_i) - 1u) : 0u;  // positions inside the list
        if (nd16 & ~comma16 & range) bad = true;                              // neither digit nor comma
        // a digit run ends at the first non-digit or at the end of the list; runs may continue in the next
        // lane's characters (for lane 31: the 16 characters staged behind the chunk)
        const uint32_t t16 = (nd16 | ~below_hi) & 0xffffu;
        uint32_t ext_t16 = 0xffffu;
        if (lane == 0) {
            const uint4 x = *reinterpret_cast<const uint4 *>(sm.mmbuf + DEC_MM_CHUNK);
            const uint32_t e_nd = pack4(nondigit_mask4(x.x)) | pack4(nondigit_mask4(x.y)) << 4 | pack4(nondigit_mask4(x.z)) << 8 |
                                  pack4(nondigit_mask4(x.w)) << 12;
            const uint32_t e_off = base + DEC_MM_CHUNK;
            const uint32_t e_hi = le > e_off ? (le - e_off < 16u ? le - e_off : 16u) : 0u;
            ext_t16 = (e_nd | ~((1u << e_hi) - 1u)) & 0xffffu;
        }
        ext_t16 = __shfl_sync(FULL_MASK, ext_t16, 0);
        uint32_t next_t16 = __shfl_down_sync(FULL_MASK, t16, 1);
        if (lane == 31) next_t16 = ext_t16;
        const uint32_t t32 = t16 | (next_t16 << 16);
        // commas owned by this lane, parsed values
        uint32_t vals[8];
        int nv = 0;
        uint32_t lsum = 0;
        const uint32_t *mmbuf32 = reinterpret_cast<const uint32_t *>(sm.mmbuf);
        for (uint32_t cm = comma16 & range; cm; cm &= cm - 1u) {
            const uint32_t i = (uint32_t)__ffs((int)cm) - 1u;
            const uint32_t run = t32 >> (i + 1u);
            const uint32_t L = run ? (uint32_t)__ffs((int)run) - 1u : 31u - i;  // digits that follow the comma
            const uint32_t q = lane * 16 + i + 1;                              // index of the first one in mmbuf
            uint32_t val;
            if (L <= 4u) {
                // up to four ASCII digits -> value, without a loop
                const uint32_t x = __funnelshift_r(mmbuf32[q >> 2], mmbuf32[(q >> 2) + 1], (q & 3u) * 8u);
                const uint32_t mask = L >= 4u ? 0xffffffffu : (1u << (8u * L)) - 1u;
                const uint32_t y = ((x & mask) - (0x30303030u & mask)) << (8u * (4u - L));  // last digit in the top byte
                val = (y >> 24) + ((y >> 16) & 0xffu) * 10u + ((y >> 8) & 0xffu) * 100u + (y & 0xffu) * 1000u;
                if (L == 0u) bad = true;
            } else {
                val = 0;
                for (uint32_t nd = 0; nd < L && nd < 10u; nd++) {
                    const uint32_t d = sm.mmbuf[q + nd];
                    val = val >= DEC_SAT / 10 ? DEC_SAT : val * 10 + (d - '0');
                }
                if (val >= DEC_SAT) val = DEC_SAT - 1;
            }
            if (nv < 8) vals[nv] = val;
            nv++;
            lsum = sat_add(lsum, val + 1);
        }
        uint32_t cnt = (uint32_t)nv;
        uint32_t incl_c = warp_inclusive_sum(cnt);
        uint32_t incl_s = warp_inclusive_sat_sum(lsum);
        if (rank_out) {
            uint32_t k = n_delta + incl_c - cnt;
            uint32_t excl_s = __shfl_up_sync(FULL_MASK, incl_s, 1);
            if (lane == 0) excl_s = 0;
            uint32_t run = sat_add(total, excl_s);
            for (int j = 0; j < nv && j < 8; j++) {
                run = sat_add(run, vals[j] + 1);
                if (k < rank_cap) rank_out[k] = run - 1;
                k++;
            }
        }
        n_delta += __shfl_sync(FULL_MASK, incl_c, 31);
        total = sat_add(total, __shfl_sync(FULL_MASK, incl_s, 31));
        __syncwarp();
    }
    bad = __any_sync(FULL_MASK, bad);
    g.n_delta = n_delta;
    g.total = total;
    __syncwarp();
    return !bad;
}

// ---------------------------------------------------------------------------------------------
// Delta list of one segment, lean form: the commas of a staged 512-byte piece are first listed (their offsets,
// compacted through a warp scan), then handled one per lane: digit run length from the staged non-digit masks,
// up to four digits converted without a loop, ranks by a plain warp scan.  Values are clamped to 65536 (the lean
// path only serves records of at most 65535 bases, so a larger skip is "beyond SEQ" either way).
// mode 0: count the deltas only (the list's ML bytes must be skipped, nothing else is needed);
// mode 1: count and sum; mode 2: also write the ranks.  `scratch` is 600 bytes of the warp's shared memory.
// Returns false (uniform) if the list is malformed or holds something unusual: the streaming path decides.
// ---------------------------------------------------------------------------------------------
__device__ bool mm_parse_list_lean(const uint8_t *mm, SegInfo &g, DecodeWarpSmem &sm, uint16_t *scratch, int mode, uint32_t *rank_out,
                                   uint32_t rank_cap) {
    const unsigned lane = lane_id();
    const uint32_t lb = g.list_begin, le = g.list_end;  // list chars are [lb, le), le is ';'
    uint16_t *cpos = scratch;        // [256] offsets of the piece's commas
    uint16_t *ndm = scratch + 256;   // [34] non-digit masks of the 33 staged 16-byte groups
    uint32_t n_delta = 0, total = 0;
    bool bad = false;
    const uint32_t *mmbuf32 = reinterpret_cast<const uint32_t *>(sm.mmbuf);
    for (uint32_t base = lb & ~15u; base < le; base += DEC_MM_CHUNK) {
        const uint32_t off = base + lane * 16;
        const uint4 v = *reinterpret_cast<const uint4 *>(mm + off);  // bytes past the string are blob padding / later fields, never used
        const uint32_t comma16 = pack4(byte_eq_mask4(v.x, 0x2c2c2c2cu)) | pack4(byte_eq_mask4(v.y, 0x2c2c2c2cu)) << 4 |
                                 pack4(byte_eq_mask4(v.z, 0x2c2c2c2cu)) << 8 | pack4(byte_eq_mask4(v.w, 0x2c2c2c2cu)) << 12;
        const uint32_t nd16 = pack4(nondigit_mask4(v.x)) | pack4(nondigit_mask4(v.y)) << 4 | pack4(nondigit_mask4(v.z)) << 8 |
                              pack4(nondigit_mask4(v.w)) << 12;
        const uint32_t lo_i = lb > off ? (lb - off < 16u ? lb - off : 16u) : 0u;
        const uint32_t hi_i = le > off ? (le - off < 16u ? le - off : 16u) : 0u;
        const uint32_t below_hi = (1u << hi_i) - 1u;
        const uint32_t range = hi_i > lo_i ? below_hi & ~((1u << lo_i) - 1u) : 0u;  // positions inside the list
        if (nd16 & ~comma16 & range) bad = true;  // neither digit nor comma
        const uint32_t commas = comma16 & range;
        const uint32_t cnt = (uint32_t)__popc(commas);
        const uint32_t incl_c = warp_inclusive_sum(cnt);
        const uint32_t n_here = __shfl_sync(FULL_MASK, incl_c, 31);
        if (mode == 0) { n_delta += n_here; continue; }
        // stage the piece (+16 bytes behind it), the non-digit masks (a run also ends where the list ends) and the comma offsets
        *reinterpret_cast<uint4 *>(sm.mmbuf + lane * 16) = v;
        ndm[lane] = (uint16_t)(nd16 | ~below_hi);
        if (lane == 0) {
            const uint4 x = *reinterpret_cast<const uint4 *>(mm + base + DEC_MM_CHUNK);
            *reinterpret_cast<uint4 *>(sm.mmbuf + DEC_MM_CHUNK) = x;
            const uint32_t e_nd = pack4(nondigit_mask4(x.x)) | pack4(nondigit_mask4(x.y)) << 4 | pack4(nondigit_mask4(x.z)) << 8 |
                                  pack4(nondigit_mask4(x.w)) << 12;
            const uint32_t e_off = base + DEC_MM_CHUNK;
            const uint32_t e_hi = le > e_off ? (le - e_off < 16u ? le - e_off : 16u) : 0u;
            ndm[32] = (uint16_t)(e_nd | ~((1u << e_hi) - 1u));
            ndm[33] = 0xffffu;
        }
        {
            uint32_t o = incl_c - cnt;
            for (uint32_t cm = commas; cm; cm &= cm - 1u) cpos[o++] = (uint16_t)(lane * 16u + (uint32_t)__ffs((int)cm) - 1u);
        }
        __syncwarp();
        for (uint32_t j0 = 0; j0 < n_here; j0 += 32) {
            const uint32_t j = j0 + lane;
            uint32_t val1 = 0;  // delta + 1
            if (j < n_here) {
                const uint32_t q = (uint32_t)cpos[j] + 1u;  // first digit, 1..512
                const uint32_t m32 = (uint32_t)ndm[q >> 4] | ((uint32_t)ndm[(q >> 4) + 1] << 16);
                const uint32_t run = m32 >> (q & 15u);
                const uint32_t L = run ? (uint32_t)__ffs((int)run) - 1u : 17u;  // digits that follow the comma
                if (L == 0u || L > 9u) bad = true;
                const uint32_t x = __funnelshift_r(mmbuf32[q >> 2], mmbuf32[(q >> 2) + 1], (q & 3u) * 8u);
                if (L <= 4u) {
                    const uint32_t mask = L >= 4u ? 0xffffffffu : (1u << (8u * L)) - 1u;
                    const uint32_t y = ((x & mask) - (0x30303030u & mask)) << (8u * (4u - L));  // last digit in the top byte
                    val1 = (y >> 24) + ((y >> 16) & 0xffu) * 10u + ((y >> 8) & 0xffu) * 100u + (y & 0xffu) * 1000u + 1u;
                } else val1 = 65536u;  // >= 10000 canonical bases skipped: never inside a record of this path
            }
            const uint32_t incl_s = warp_inclusive_sum(val1);
            if (mode == 2 && j < n_here) {
                const uint32_t k = n_delta + j, rk = total + incl_s - 1u;
                if (k < rank_cap) rank_out[k] = rk < DEC_SAT ? rk : DEC_SAT;
            }
            total += __shfl_sync(FULL_MASK, incl_s, 31);
            if (total > DEC_SAT) total = DEC_SAT;
        }
        n_delta += n_here;
        __syncwarp();
    }
    g.n_delta = n_delta;
    g.total = total;
    __syncwarp();
    return !__any_sync(FULL_MASK, bad);
}

// ---------------------------------------------------------------------------------------------
// get_mod_poss_on_ref with implicit canonical calls (blockjoin.c:605-792 with seqi given: 666-700, 727-761): a listed
// cytosine outside a CpG makes the reference fill in every unlisted CpG of the aligned stretches as unmethylated.  The
// walk over CIGAR operations and kept mods (read offset, category; trig_p / trig_cat, n_mods of them, in ascending or
// descending read order) stays sequential, executed by the whole warp uniformly, statement by statement as in
// decode_generic(); the CpG scan of the stretch between two mods is done by the 32 lanes.  Writes the calls to
// opos / ocat (cap slots), *n_out = calls produced; returns RS_KEPT | RS_OVERFLOW | RS_UNSORTED | RS_FATAL_CIGAR bits.
// ---------------------------------------------------------------------------------------------
__device__ uint32_t implicit_walk(const uint32_t *cigar, uint32_t n_cigar, const uint8_t *seq, uint32_t len, uint32_t qs, bool rev,
                                  const uint32_t *trig_p, const uint8_t *trig_cat, uint32_t n_mods, bool descending, uint32_t *opos,
                                  uint8_t *ocat, uint32_t cap, uint32_t *n_out) {
    const unsigned lane = lane_id();
    const int cg = rev ? -1 : 0;
    bool fatal = false;
    uint32_t n = 0, last = 0;
    bool uns = false;
    auto trig = [&](uint32_t k, uint32_t *tp, uint32_t *tc) {
        const uint32_t ti = descending ? n_mods - 1u - k : k;  // ascending read offsets
        *tp = trig_p[ti]; *tc = trig_cat[ti];
    };
    auto push = [&](uint32_t pp, uint32_t c) {  // CallSink::push
        if (n > 0 && n <= cap && pp <= last) uns = true;
        if (n < cap) { if (lane == 0) { opos[n] = pp; ocat[n] = (uint8_t)c; } last = pp; }
        n++;
    };
    auto fill = [&](uint32_t from, uint32_t until, uint32_t i_ref_, int32_t offset_) {  // gen_implicit_fill, 32 positions per step
        for (uint32_t b0 = from; b0 < until; b0 += 32) {
            const uint32_t t = b0 + lane;
            const bool hit = t < until && t < len - 1u && seq_nib(seq, t) == 2u && seq_nib(seq, t + 1u) == 4u;
            unsigned hm = __ballot_sync(FULL_MASK, hit);
            if (!hm) continue;
            // positions ascend inside a stretch: only its first CpG can sit on the call pushed last ("implicit but not pushing")
            const uint32_t p_first = i_ref_ + b0 + (uint32_t)__ffs((int)hm) - 1u + (uint32_t)offset_;
            if (n > 0 && n <= cap && last == p_first) hm &= hm - 1u;
            if (!hm) continue;
            const uint32_t p_lo = i_ref_ + b0 + (uint32_t)__ffs((int)hm) - 1u + (uint32_t)offset_;
            if (n > 0 && n <= cap && p_lo <= last) uns = true;
            if ((hm >> lane) & 1u) {
                const uint32_t idx = n + (uint32_t)__popc(hm & ((1u << lane) - 1u));
                if (idx < cap) { opos[idx] = i_ref_ + t + (uint32_t)offset_; ocat[idx] = 1; }
            }
            const uint32_t cnt = (uint32_t)__popc(hm);
            if (n + cnt - 1u < cap) last = i_ref_ + b0 + (31u - (uint32_t)__clz((int)hm)) + (uint32_t)offset_;
            n += cnt;
            __syncwarp();
        }
    };
    uint32_t i_read2 = 0, i_ref2 = qs, it = 0, next, nq;
    trig(0, &next, &nq);
    uint32_t ic = 0;
    if ((cigar[0] & 15u) == 4u) {
        i_read2 = cigar[0] >> 4;
        while (next < i_read2) {
            it++;
            if (it < n_mods) trig(it, &next, &nq); else break;
        }
        if (next == i_read2) {
            push(i_ref2 + (uint32_t)cg, nq);
            it++;
            if (it < n_mods) trig(it, &next, &nq);
        }
        i_ref2 -= cigar[0] >> 4;
        ic = 1;
    }
    int32_t off2 = 0;
    for (; ic < n_cigar; ic++) {
        const uint32_t op = cigar[ic] & 15u, L = cigar[ic] >> 4;
        if (op <= 1u) {
            uint32_t pos_canonical = i_read2;
            while (i_read2 + L >= next) {
                if (op == 0u && next != 0xffffffffu) {
                    const uint32_t until = next - 1u < i_read2 + L ? next - 1u : i_read2 + L;
                    fill(pos_canonical, until, i_ref2, off2);
                    const uint32_t pt = i_ref2 + next + (uint32_t)cg + (uint32_t)off2;
                    if (n > 0 && n <= cap && last == pt) { if (lane == 0) ocat[n - 1u] = (uint8_t)nq; }  // CallSink::last_is / set_last_cat
                    else push(pt, nq);
                    pos_canonical = cg == 0 ? next + 1u : next + 2u;
                }
                it++;
                if (it >= n_mods) { next = 0xffffffffu; break; }
                trig(it, &next, &nq);
            }
            if (op == 0u) {
                fill(pos_canonical, i_read2 + L, i_ref2, off2);
                i_read2 += L;
            } else { i_read2 += L; off2 -= (int32_t)L; }
        } else if (op == 2u) off2 += (int32_t)L;
        else if (op == 3u || op == 4u) break;
        else { fatal = true; break; }
    }
    __syncwarp();
    *n_out = n;
    if (fatal) return RS_FATAL_CIGAR;
    return RS_KEPT | (n > cap ? RS_OVERFLOW : 0u) | (uns ? RS_UNSORTED : 0u);
}

// ---------------------------------------------------------------------------------------------
// The warp-parallel fast path.  Returns status bits; n_calls_out receives the number of calls.
// ---------------------------------------------------------------------------------------------
__device__ uint32_t decode_fast(const DecodeParams &P, const ReadRec &R, DecodeWarpSmem &sm, uint32_t *n_calls_out,
                                bool *need_generic) {
    const unsigned lane = lane_id();
    const uint8_t *blob = P.blob;
    const uint32_t *cigar = reinterpret_cast<const uint32_t *>(blob + (size_t)R.cigar_off * 16);
    const uint8_t *seq = blob + (size_t)R.seq_off * 16;
    const uint8_t *mm = blob + (size_t)R.mm_off * 16;
    const uint8_t *ml = blob + (size_t)R.ml_off * 16;
    const uint32_t len = R.l_qseq;
    const bool rev = (R.flags & 16u) != 0;
    const bool has_ml = (R.flags & RF_HAS_ML) != 0;
    uint32_t *rank = P.tmp_rank + R.calls_off;
    uint32_t *mpos = P.tmp_mpos + R.calls_off;
    uint8_t *mcat = P.tmp_mcat + R.calls_off;
    uint32_t *opos = P.calls_pos + R.calls_off;
    uint8_t *ocat = P.calls_cat + R.calls_off;
    const uint32_t cap = R.calls_cap;
    uint32_t status = 0;
    *n_calls_out = 0;
    *need_generic = false;

    // ---- MM structure ----
    int n_seg = 0;
    bool mm_error = false;
    if (!(R.flags & RF_HAS_MM)) n_seg = 0;
    else if (R.flags & RF_MALFORMED) mm_error = true;
    else {
        n_seg = mm_scan_segments(mm, R.mm_len, sm);
        if (n_seg == -2) { *need_generic = true; return 0; }
        if (n_seg < 0) { mm_error = true; n_seg = 0; }
    }
    int rel = -1;
    if (!mm_error) {
        int n_rel = 0;
        for (int s = 0; s < n_seg; s++) {
            const SegInfo &g = sm.seg[s];
            if (g.canon == 2 && g.m_idx >= 0) { n_rel += g.m_count; rel = s; }
        }
        if (n_rel > 1) { *need_generic = true; return 0; }
    }
    // ---- delta lists ----
    uint32_t ml_need = 0;
    if (!mm_error) {
        for (int s = 0; s < n_seg; s++) {
            SegInfo &g = sm.seg[s];
            const bool want_ranks = (s == rel);
            // forward strand: other segments only need their comma count (for ML offsets); counting and
            // parsing share one pass either way
            bool ok = mm_parse_list(mm, g, sm, want_ranks ? rank : nullptr, cap);
            if (!ok) { mm_error = true; break; }
            if (lane == 0) sm.seg[s].ml_base = ml_need;
            ml_need += g.n_delta * g.n_codes;
            if (has_ml && ml_need > R.ml_len) { mm_error = true; break; }
            if (want_ranks && g.n_delta > cap) { *need_generic = true; return 0; }  // cannot happen: cap >= ML length
        }
        if (!mm_error && has_ml && ml_need != R.ml_len) mm_error = true;
        __syncwarp();
    }

    // ---- SEQ scan: select the listed canonical bases of the relevant segment ----
    // Pass 1 streams SEQ (two 512-byte tiles per step, next step prefetched): every lane counts the canonical
    // bases of its 32-base chunk and a warp scan gives the rank of the chunk's first one, kept in shared
    // memory.  Pass 2 is dense over the listed bases: one lane per base binary-searches the chunk that holds
    // its rank, selects the base inside the chunk, checks the CpG context and turns the ML byte into a
    // category; ML reads and the stores are coalesced in list order.  Reverse-strand records are scanned
    // from the right end of SEQ so ranks count from the read's own 5' end (htslib walks the delta list
    // backwards instead).
    uint32_t n_mods = 0, mbase = 0, n_listed = 0;
    bool has_implicit = false;
    const bool do_select = !mm_error && rel >= 0 && sm.seg[rel].n_delta > 0;
    // reverse strand: "MM tag refers to bases beyond sequence length" iff count(complement base) < sum(delta+1)
    uint32_t need_c = 0;
    const uint32_t n_bytes = (len + 1) >> 1;
    if (rev && !mm_error && len > 0) {
        for (int s = 0; s < n_seg; s++) {
            const SegInfo &g = sm.seg[s];
            if (g.n_codes == 0) continue;
            if (g.canon == 2) need_c = g.total > need_c ? g.total : need_c;
            else if (g.total > 0 && g.total > count_base(seq, n_bytes, comp_code(g.canon))) mm_error = true;
        }
    }
    if (!mm_error && len > 0 && (do_select || need_c > 0)) {
        const uint32_t pat = (rev ? 4u : 2u) * 0x11111111u;  // C on the read strand is G in SEQ for reverse alignments
        const uint32_t n_targets = do_select ? sm.seg[rel].n_delta : 0;
        const uint32_t ml_base = do_select ? sm.seg[rel].ml_base : 0;
        const uint32_t stride = do_select ? sm.seg[rel].n_codes : 0;
        const uint32_t m_idx = do_select ? (uint32_t)sm.seg[rel].m_idx : 0;
        const int n_tiles = (int)((n_bytes + 511) / 512);
        uint32_t run = 0;        // canonical bases seen so far, in scan order
        uint32_t tcur = 0;       // listed bases with rank < run (list order == scan order on both strands)
        bool drop_first = false, drop_last = false, implicit = false;
        // SEQ is walked in sections of up to DEC_FC chunks (16 bytes = 32 bases each), in scan order
        for (int tile0 = 0; tile0 < n_tiles; tile0 += DEC_FC / 32) {
            if (tcur >= n_targets && run >= need_c) break;  // everything listed was found (and counted, reverse)
            const int sec_tiles = n_tiles - tile0 < DEC_FC / 32 ? n_tiles - tile0 : DEC_FC / 32;
            const int n_steps = (sec_tiles + DEC_T - 1) / DEC_T;
            const uint32_t run0 = run;
            // ---- pass 1: stream the section, rank of every chunk's first match -> sm.first[] ----
            uint4 cur[DEC_T], nxt[DEC_T];
#pragma unroll
            for (int t = 0; t < DEC_T; t++) {
                const int ts = tile0 + t;  // tile in scan order
                const int tile = rev ? n_tiles - 1 - ts : ts;
                const uint32_t byte_off = (uint32_t)tile * 512u + lane * 16u;
                cur[t] = make_uint4(0, 0, 0, 0);
                if (t < sec_tiles && byte_off < n_bytes) cur[t] = *reinterpret_cast<const uint4 *>(seq + byte_off);
            }
            for (int step = 0; step < n_steps; step++) {
#pragma unroll
                for (int t = 0; t < DEC_T; t++) {  // prefetch the next step
                    const int tl = (step + 1) * DEC_T + t;  // tile inside the section
                    const int tile = rev ? n_tiles - 1 - (tile0 + tl) : tile0 + tl;
                    const uint32_t byte_off = (uint32_t)tile * 512u + lane * 16u;
                    nxt[t] = make_uint4(0, 0, 0, 0);
                    if (tl < sec_tiles && byte_off < n_bytes) nxt[t] = *reinterpret_cast<const uint4 *>(seq + byte_off);
                }
                uint32_t C[DEC_T], packed = 0;
#pragma unroll
                for (int t = 0; t < DEC_T; t++) {
                    C[t] = (uint32_t)(__popc(nib_eq_flags(cur[t].x, pat)) + __popc(nib_eq_flags(cur[t].y, pat)) +
                                      __popc(nib_eq_flags(cur[t].z, pat)) + __popc(nib_eq_flags(cur[t].w, pat)));
                    packed |= C[t] << (16 * t);  // a tile holds at most 1024 matches: 16 bits per tile
                }
                const uint32_t incl = warp_inclusive_sum(packed);
                const uint32_t tot = __shfl_sync(FULL_MASK, incl, 31);
                uint32_t before = 0;
#pragma unroll
                for (int t = 0; t < DEC_T; t++) {
                    const uint32_t tile_total = (tot >> (16 * t)) & 0xffffu;
                    const uint32_t incl_t = (incl >> (16 * t)) & 0xffffu;
                    // rank of this lane's first match; the chunk's index in scan order
                    const uint32_t fr = run + before + (rev ? tile_total - incl_t : incl_t - C[t]);
                    const uint32_t ci = (uint32_t)(step * DEC_T + t) * 32u + (rev ? 31u - lane : lane);
                    if (step * DEC_T + t < sec_tiles) sm.first[ci] = fr;
                    before += tile_total;
                }
                run += before;
#pragma unroll
                for (int t = 0; t < DEC_T; t++) cur[t] = nxt[t];
            }
            const uint32_t n_ch = (uint32_t)sec_tiles * 32u;
            if (lane == 0) sm.first[n_ch] = run;
            __syncwarp();
            // ---- pass 2: the listed bases whose rank falls into the section, 32 at a time ----
            while (tcur < n_targets) {
                const uint32_t k = tcur + lane;
                const uint32_t r = k < n_targets ? rank[k] : 0xffffffffu;
                const bool mine = r < run;  // ranks ascend strictly and every rank below run0 is consumed
                const unsigned act = __ballot_sync(FULL_MASK, mine);
                if (act == 0) break;
                if (mine) {
                    // last chunk whose first rank is <= r (empty chunks share the rank of their successor)
                    uint32_t lo = 0, hi = n_ch;  // first[lo] <= r < first[hi]
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (sm.first[mid] <= r) lo = mid; else hi = mid;
                    }
                    const uint32_t f0 = sm.first[lo], cnt = sm.first[lo + 1] - f0;
                    const uint32_t ts = (uint32_t)tile0 + (lo >> 5);
                    const uint32_t tile = rev ? (uint32_t)n_tiles - 1u - ts : ts;
                    const uint32_t chunk = tile * 32u + (rev ? 31u - (lo & 31u) : lo & 31u);
                    uint32_t n = rev ? cnt - 1u - (r - f0) : r - f0;  // index in base order inside the chunk
                    const uint4 v = *reinterpret_cast<const uint4 *>(seq + (size_t)chunk * 16);
                    const uint32_t g0 = nib_eq_flags(v.x, pat), g1 = nib_eq_flags(v.y, pat), g2 = nib_eq_flags(v.z, pat),
                                   g3 = nib_eq_flags(v.w, pat);
                    const uint32_t c0 = (uint32_t)__popc(g0), c1 = (uint32_t)__popc(g1), c2 = (uint32_t)__popc(g2);
                    uint32_t wj = 0, f = g0;
                    if (n >= c0) {
                        n -= c0; wj = 1; f = g1;
                        if (n >= c1) {
                            n -= c1; wj = 2; f = g2;
                            if (n >= c2) { n -= c2; wj = 3; f = g3; }
                        }
                    }
                    const uint32_t p = chunk * 32u + wj * 8u + select_base_in_word(f, n);
                    // blockjoin.c:846-858: C must be followed by G; on reverse alignments SEQ shows the G, preceded by C
                    const uint32_t slot = rev ? cap - 1u - k : k;
                    if (p > 0 && p < len - 1) {
                        const bool ok = rev ? seq_nib(seq, p - 1) == 2u : seq_nib(seq, p + 1) == 4u;
                        if (ok) {
                            const uint32_t q = has_ml ? ml[ml_base + k * stride + m_idx] : 255u;
                            mpos[slot] = p;
                            mcat[slot] = (uint8_t)(q < P.lo ? 1 : (q >= P.hi ? 0 : 2));  // blockjoin.c:876-878
                        } else { implicit = true; mpos[slot] = 0xffffffffu; }  // (a hole: squeezed out below if the record has implicit calls)
                    } else {
                        mpos[slot] = 0xffffffffu;
                        if (k == 0) drop_first = true; else drop_last = true;
                    }
                }
                tcur += (uint32_t)__popc(act);
                if (act != FULL_MASK) break;
            }
            __syncwarp();
            (void)run0;
        }
        if (rev && run < need_c) mm_error = true;
        has_implicit = __any_sync(FULL_MASK, implicit);
        // only the first and the last base of SEQ can be listed and silently dropped (0 < pos < len-1)
        const uint32_t d0 = __any_sync(FULL_MASK, drop_first) ? 1u : 0u, d1 = __any_sync(FULL_MASK, drop_last) ? 1u : 0u;
        n_mods = tcur - d0 - d1;
        mbase = rev ? cap - tcur + d1 : d0;  // mods occupy tmp[mbase, mbase+n_mods), ascending SEQ position
        n_listed = tcur;
    }
    if (mm_error) { status |= RS_MM_ERROR; n_mods = 0; has_implicit = false; }
    __syncwarp();
    if (has_implicit) {
        // Implicit canonical calls: the listed bases that are not CpGs left holes among the kept mods; squeeze them out
        // (ascending SEQ position either way: slot k on forward, cap-1-k on reversed alignments), then the sequential
        // walk with the lanes scanning the stretches (implicit_walk).
        const uint32_t lbase = rev ? cap - n_listed : 0u;
        uint32_t nk = 0;
        for (uint32_t i0 = 0; i0 < n_listed; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t pp = 0xffffffffu;
            uint8_t cc = 0;
            if (i < n_listed) { pp = mpos[lbase + i]; cc = mcat[lbase + i]; }
            const unsigned vm = __ballot_sync(FULL_MASK, pp != 0xffffffffu);
            __syncwarp();
            if (pp != 0xffffffffu) {
                const uint32_t d = nk + (uint32_t)__popc(vm & ((1u << lane) - 1u));
                mpos[lbase + d] = pp; mcat[lbase + d] = cc;
            }
            nk += (uint32_t)__popc(vm);
            __syncwarp();
        }
        status |= RS_HAS_IMPLICIT;
        if (R.n_cigar == 0 || nk == 0) return status;  // get_mod_poss_on_ref returns 0: record dropped
        uint32_t n = 0;
        const uint32_t wst = implicit_walk(cigar, R.n_cigar, seq, len, R.pos, rev, mpos + lbase, mcat + lbase, nk, false, opos, ocat, cap, &n);
        *n_calls_out = n;
        return status | wst;
    }

    // ---- CIGAR walk ----
    const uint32_t n_cigar = R.n_cigar;
    if (n_cigar == 0 || n_mods == 0) return status;  // get_mod_poss_on_ref returns 0: record dropped
    const int cg = rev ? -1 : 0;
    const uint32_t qs = R.pos;
    uint32_t clip = 0, j0 = 0;
    {
        uint32_t c0 = cigar[0];
        if ((c0 & 15u) == 4u) { clip = c0 >> 4; j0 = 1; }
    }
    uint32_t n_out = 0;
    uint32_t last_pos = 0;
    bool have_last = false, unsorted = false;
    // leading soft clip: silently consume mods inside it, emit a mod sitting exactly at the clip edge
    uint32_t it0 = 0;
    if (j0) {
        // first index with t >= clip (mods ascending)
        uint32_t lo_i = 0, hi_i = n_mods;
        while (lo_i < hi_i) {
            uint32_t mid = (lo_i + hi_i) >> 1;
            if (mpos[mbase + mid] < clip) lo_i = mid + 1; else hi_i = mid;
        }
        it0 = lo_i;
        if (it0 < n_mods && mpos[mbase + it0] == clip) {
            if (lane == 0 && n_out < cap) { opos[n_out] = qs + (uint32_t)cg; ocat[n_out] = mcat[mbase + it0]; }
            last_pos = qs + (uint32_t)cg;
            have_last = true;
            n_out = 1;
            it0++;
        }
    }
    bool stale = j0 && it0 >= n_mods;  // every trigger was consumed by the clip: the last one lingers
    const uint32_t stale_t = mpos[mbase + n_mods - 1];
    const uint8_t stale_cat = mcat[mbase + n_mods - 1];
    const uint32_t i_ref = qs - clip;
    uint32_t i_read = clip;
    int32_t offset = 0;
    uint32_t tcur = it0;
    uint32_t n_mi = 0;
    bool stopped = false, fatal = false, first_mi_seen = false;
    __syncwarp();
    for (uint32_t cbase = j0; !stopped; cbase += 32) {
        const bool more = cbase < n_cigar;
        uint32_t op = 0xf, L = 0;
        bool in = false;
        if (more) {
            uint32_t idx = cbase + lane;
            in = idx < n_cigar;
            if (in) { uint32_t c = cigar[idx]; op = c & 15u; L = c >> 4; }
        }
        // first stopping op in this group of 32
        unsigned stopm = __ballot_sync(FULL_MASK, in && op >= 3u);
        int stop_lane = stopm ? __ffs(stopm) - 1 : 32;
        if (stopm) {
            uint32_t sop = __shfl_sync(FULL_MASK, op, stop_lane);
            if (sop >= 5u) fatal = true;
            stopped = true;
        }
        if (!more) stopped = true;
        const bool act = in && (int)lane < stop_lane;
        const bool isMI = act && op <= 1u;
        uint32_t adv = isMI ? L : 0u;
        int32_t doff = !act ? 0 : (op == 2u ? (int32_t)L : (op == 1u ? -(int32_t)L : 0));
        uint32_t incl_adv = warp_inclusive_sum(adv);
        int32_t incl_off = warp_inclusive_sum(doff);
        unsigned mim = __ballot_sync(FULL_MASK, isMI);
        if (isMI) {
            uint32_t slot = n_mi + __popc(mim & ((1u << lane) - 1u));
            sm.mi_end[slot] = i_read + incl_adv;
            sm.mi_off[slot] = op == 0u ? offset + (incl_off - doff) : MI_DROP;
        }
        n_mi += __popc(mim);
        i_read += __shfl_sync(FULL_MASK, incl_adv, 31);
        offset += __shfl_sync(FULL_MASK, incl_off, 31);
        __syncwarp();
        // flush when the stage is nearly full or the walk is over
        if (n_mi + 32 > DEC_MI_CAP || stopped) {
            if (n_mi > 0) {
                if (stale && !first_mi_seen) {
                    // blockjoin.c:663-665 with a lingering trigger: the first M/I op handles it once
                    int32_t o = sm.mi_off[0];
                    if (o != MI_DROP) {
                        uint32_t pos = i_ref + stale_t + (uint32_t)cg + (uint32_t)o;
                        if (have_last && last_pos == pos) {
                            if (lane == 0) ocat[n_out - 1] = stale_cat;
                        } else {
                            if (lane == 0 && n_out < cap) { opos[n_out] = pos; ocat[n_out] = stale_cat; }
                            if (have_last && pos < last_pos) unsorted = true;
                            n_out++;
                            last_pos = pos;
                            have_last = true;
                        }
                    }
                    stale = false;
                }
                first_mi_seen = true;
                const uint32_t chunk_last_end = sm.mi_end[n_mi - 1];
                while (tcur < n_mods) {
                    uint32_t it = tcur + lane;
                    uint32_t t = 0;
                    bool mine = false;
                    if (it < n_mods) { t = mpos[mbase + it]; mine = t <= chunk_last_end; }
                    unsigned am = __ballot_sync(FULL_MASK, mine);
                    if (am == 0) break;
                    bool emit = false;
                    uint32_t pos = 0;
                    uint8_t cat = 0;
                    if (mine) {
                        uint32_t lo_j = 0, hi_j = n_mi - 1;  // first op with end >= t
                        while (lo_j < hi_j) {
                            uint32_t mid = (lo_j + hi_j) >> 1;
                            if (sm.mi_end[mid] >= t) hi_j = mid; else lo_j = mid + 1;
                        }
                        int32_t o = sm.mi_off[lo_j];
                        if (o != MI_DROP) {
                            emit = true;
                            pos = i_ref + t + (uint32_t)cg + (uint32_t)o;
                            cat = mcat[mbase + it];
                        }
                    }
                    // de-duplicate against the previously emitted position (blockjoin.c:704-709)
                    unsigned em = __ballot_sync(FULL_MASK, emit);
                    unsigned below = em & ((1u << lane) - 1u);
                    int prev_lane = below ? 31 - __clz((int)below) : -1;
                    uint32_t prev_pos = __shfl_sync(FULL_MASK, pos, prev_lane < 0 ? 0 : prev_lane);
                    bool has_prev = prev_lane >= 0 || have_last;
                    if (prev_lane < 0) prev_pos = last_pos;
                    bool head = emit && !(has_prev && prev_pos == pos);
                    if (head && has_prev && pos < prev_pos) unsorted = true;
                    unsigned hm = __ballot_sync(FULL_MASK, head);
                    // index of the run this lane belongs to
                    uint32_t heads_incl = __popc(hm & ((2u << lane) - 1u));
                    uint32_t slot = n_out + heads_incl - 1;  // for a continuation of the carried run heads_incl==0 -> n_out-1
                    unsigned above = em & ~((2u << lane) - 1u);
                    int next_lane = above ? __ffs(above) - 1 : -1;
                    uint32_t next_pos = __shfl_sync(FULL_MASK, pos, next_lane < 0 ? 0 : next_lane);
                    bool last_of_run = emit && (next_lane < 0 || next_pos != pos);
                    if (head && slot < cap) opos[slot] = pos;
                    if (last_of_run && slot < cap) ocat[slot] = cat;
                    n_out += __popc(hm);
                    if (em) {
                        int ll = 31 - __clz((int)em);
                        last_pos = __shfl_sync(FULL_MASK, pos, ll);
                        have_last = true;
                    }
                    tcur += __popc(am);
                    __syncwarp();
                    if (__popc(am) < 32) break;
                }
            }
            n_mi = 0;
        }
    }
    unsorted = __any_sync(FULL_MASK, unsorted);
    if (fatal) return status | RS_FATAL_CIGAR;
    if (n_out > cap) status |= RS_OVERFLOW;
    if (unsorted) status |= RS_UNSORTED;
    *n_calls_out = n_out;
    return status | RS_KEPT;
}

// ---------------------------------------------------------------------------------------------
// The lean path: records of at most LEAN_MAXLEN bases with fewer than LEAN_MI M/I operations in front of the
// first stopping operation, one C+m stream, no implicit calls — i.e. ordinary reads.  Both per-record tables
// (canonical-base ranks per 16-byte SEQ chunk, M/I operations) stay in shared memory as 16-bit values for the
// whole record, so every listed base is carried from its MM rank to its reference position in one pass:
// rank -> SEQ chunk (branch-free search) -> offset in the read -> CpG test -> ML category -> M/I operation
// (branch-free search) -> reference position -> ordered, de-duplicated store.  Nothing goes through the scratch
// arrays except the ranks themselves.  Returns false when the record needs the streaming / general path
// (nothing final has been written then).
// ---------------------------------------------------------------------------------------------
__device__ bool decode_lean(const DecodeParams &P, const ReadRec &R, DecodeWarpSmem &sm, uint32_t *status_out,
                            uint32_t *n_calls_out, uint32_t *rlen_out) {
    const unsigned lane = lane_id();
    if (!(R.flags & RF_HAS_MM) || (R.flags & RF_MALFORMED)) return false;
    const uint32_t len = R.l_qseq, n_cigar = R.n_cigar;
    if (len < 2 || len > (uint32_t)LEAN_MAXLEN || n_cigar == 0) return false;
    const uint8_t *blob = P.blob;
    const uint32_t *cigar = reinterpret_cast<const uint32_t *>(blob + (size_t)R.cigar_off * 16);
    const uint8_t *seq = blob + (size_t)R.seq_off * 16;
    const uint8_t *mm = blob + (size_t)R.mm_off * 16;
    const uint8_t *ml = blob + (size_t)R.ml_off * 16;
    const bool rev = (R.flags & 16u) != 0;
    const bool has_ml = (R.flags & RF_HAS_ML) != 0;
    uint32_t *rank = P.tmp_rank + R.calls_off;
    uint32_t *opos = P.calls_pos + R.calls_off;
    uint8_t *ocat = P.calls_cat + R.calls_off;
    uint32_t *trig_p = P.tmp_mpos + R.calls_off;
    uint8_t *trig_cat = P.tmp_mcat + R.calls_off;
    const uint32_t cap = R.calls_cap;

    // ---- MM structure and delta lists (ranks of the C+m list -> rank[]) ----
    const int n_seg = mm_scan_segments(mm, R.mm_len, sm);
    if (n_seg <= 0) return false;
    int rel = -1, n_rel = 0;
    bool other_canon = false;
    for (int s = 0; s < n_seg; s++) {
        const SegInfo &g = sm.seg[s];
        if (g.canon == 2 && g.m_idx >= 0) { n_rel += g.m_count; rel = s; }
        else if (g.canon != 2 && g.n_codes) other_canon = true;
    }
    if (n_rel != 1 || (rev && other_canon)) return false;
    uint32_t ml_need = 0, need_c = 0;
    for (int s = 0; s < n_seg; s++) {
        SegInfo &g = sm.seg[s];
        // (a list other than the C+m one only matters for its length, and on reversed alignments for its sum)
        const int mode = s == rel ? 2 : (rev && g.n_codes ? 1 : 0);
        if (!mm_parse_list_lean(mm, g, sm, sm.l_first, mode, rank, cap)) return false;
        if (lane == 0) sm.seg[s].ml_base = ml_need;
        ml_need += g.n_delta * g.n_codes;
        if (has_ml && ml_need > R.ml_len) return false;
        if (g.canon == 2 && g.n_codes && g.total > need_c) need_c = g.total;
    }
    if (has_ml && ml_need != R.ml_len) return false;
    __syncwarp();
    const uint32_t n_targets = sm.seg[rel].n_delta;
    if (n_targets == 0 || n_targets > cap) return false;
    const uint32_t ml_base = sm.seg[rel].ml_base, stride = sm.seg[rel].n_codes, m_idx = (uint32_t)sm.seg[rel].m_idx;

    // ---- tables start out as "greater than any position" ----
    {
        const uint4 ones = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        uint4 *f4 = reinterpret_cast<uint4 *>(sm.l_first);
        for (uint32_t i = lane; i < (LEAN_CH + 8) / 8; i += 32) f4[i] = ones;
        uint4 *e4 = reinterpret_cast<uint4 *>(sm.l_end);
        for (uint32_t i = lane; i < LEAN_MI / 8; i += 32) e4[i] = ones;
    }
    __syncwarp();

    // ---- CIGAR, four operations per lane: reference span, M/I table up to the first stopping operation ----
    uint32_t clip = 0, j0 = 0;
    {
        const uint32_t c0 = cigar[0];
        if ((c0 & 15u) == 4u) { clip = c0 >> 4; j0 = 1; }
    }
    uint32_t rlen = 0, n_mi = 0, i_read = clip;
    int32_t offset = 0;
    bool stopped = false, fatal = false, bad = false;
    const uint4 *cig4 = reinterpret_cast<const uint4 *>(cigar);
    for (uint32_t base = 0; base < n_cigar; base += 128) {
        const uint32_t idx0 = base + lane * 4;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (idx0 < n_cigar) v = cig4[idx0 >> 2];  // the field is zero padded to 16 bytes
        const uint32_t c[4] = {v.x, v.y, v.z, v.w};
        uint32_t my_stop = 4, my_stop_op = 0;
#pragma unroll
        for (int i = 3; i >= 0; i--) {
            const uint32_t op = c[i] & 15u;
            const bool valid = idx0 + i < n_cigar;
            if (valid && ((0x18du >> op) & 1u)) rlen += c[i] >> 4;  // M, D, N, =, X consume the reference
            if (valid && idx0 + i >= j0 && op >= 3u) { my_stop = (uint32_t)i; my_stop_op = op; }
        }
        if (stopped) continue;
        const unsigned stops = __ballot_sync(FULL_MASK, my_stop < 4u);
        const int stop_lane = stops ? __ffs((int)stops) - 1 : 32;
        const uint32_t lim = (int)lane < stop_lane ? 4u : ((int)lane == stop_lane ? my_stop : 0u);
        uint32_t adv = 0, cnt = 0, l_adv[4];
        int32_t doff = 0, l_doff[4];
        bool l_mi[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t op = c[i] & 15u, L = c[i] >> 4;
            const bool act = (uint32_t)i < lim && idx0 + i < n_cigar && idx0 + i >= j0;
            l_doff[i] = doff;
            l_mi[i] = act && op <= 1u;
            if (l_mi[i]) { adv += L; cnt++; if (L > 0xffffu) bad = true; }
            l_adv[i] = adv;
            if (act) doff += op == 2u ? (int32_t)L : (op == 1u ? -(int32_t)L : 0);
        }
        const uint32_t packed = adv | (cnt << 24);  // per lane: adv <= 4 * 65535, over the warp < 2^24 (else `bad`)
        const uint32_t incl = warp_inclusive_sum(packed);
        const int32_t incl_d = warp_inclusive_sum(doff);
        const uint32_t ex_adv = (incl - packed) & 0xffffffu, ex_cnt = (incl - packed) >> 24;
        const int32_t ex_d = incl_d - doff;
        uint32_t slot = n_mi + ex_cnt;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (l_mi[i]) {
                const uint32_t e = i_read + ex_adv + l_adv[i];
                const int32_t o = offset + ex_d + l_doff[i];
                const bool is_m = (c[i] & 15u) == 0u;
                if (e > 0xffffu || (is_m && (o < -32767 || o > 32767))) bad = true;
                if (slot < (uint32_t)LEAN_MI) { sm.l_end[slot] = (uint16_t)e; sm.l_off[slot] = is_m ? (int16_t)o : LEAN_DROP; }
                slot++;
            }
        }
        const uint32_t tot = __shfl_sync(FULL_MASK, incl, 31);
        n_mi += tot >> 24;
        i_read += tot & 0xffffffu;
        offset += __shfl_sync(FULL_MASK, incl_d, 31);
        if (stops) {
            stopped = true;
            fatal = __shfl_sync(FULL_MASK, my_stop_op, stop_lane) >= 5u;
        }
    }
    rlen = warp_sum(rlen);
    if (__any_sync(FULL_MASK, bad) || fatal || n_mi >= (uint32_t)LEAN_MI) return false;  // (a fatal operation is reported by the streaming path)
    *rlen_out = rlen;
    __syncwarp();

    // ---- SEQ pass: canonical bases per chunk -> l_first[] (scan order: from the read's own 5' end) ----
    // Super-tiles of 2 KB (128 chunks, 4096 bases): every lane owns four consecutive chunks, so one warp scan and
    // one 8-byte store serve 128 chunks.  Reversed alignments walk SEQ from its end (chunk g <-> scan index
    // n_ch - 1 - g): ranks then count from the read's own 5' end, as the MM list does.
    const uint32_t n_bytes = (len + 1) >> 1;
    const uint32_t n_st = (n_bytes + 2047u) >> 11;
    const uint32_t n_ch = n_st * 128u;
    const uint32_t tab_lo = rev ? 0u : 0x00010000u, tab_hi = rev ? 0x00000001u : 0u;  // count G (4) on reversed alignments, C (2) otherwise
    uint32_t total = 0;
    {
        const uint4 zero = make_uint4(0, 0, 0, 0);
        uint4 cur[4];
        auto load_st = [&](uint32_t st) {  // scan-order super-tile st: this lane's four chunks, in memory order
            const uint32_t g0 = rev ? (n_st - 1u - st) * 128u + 4u * (31u - lane) : st * 128u + 4u * lane;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t byte_off = (g0 + (uint32_t)i) * 16u;
                cur[i] = st < n_st && byte_off < n_bytes ? *reinterpret_cast<const uint4 *>(seq + byte_off) : zero;
            }
        };
        load_st(0);
        for (uint32_t st = 0; st < n_st; st++) {
            const uint32_t K0 = count_code_in_chunk(cur[0], tab_lo, tab_hi), K1 = count_code_in_chunk(cur[1], tab_lo, tab_hi),
                           K2 = count_code_in_chunk(cur[2], tab_lo, tab_hi), K3 = count_code_in_chunk(cur[3], tab_lo, tab_hi);
            const uint32_t C[4] = {rev ? K3 : K0, rev ? K2 : K1, rev ? K1 : K2, rev ? K0 : K3};  // scan order inside the lane
            load_st(st + 1);  // in flight during the scan
            const uint32_t mine = C[0] + C[1] + C[2] + C[3];
            const uint32_t incl = warp_inclusive_sum(mine);
            const uint32_t f0 = total + incl - mine, f1 = f0 + C[0], f2 = f1 + C[1], f3 = f2 + C[2];
            *reinterpret_cast<uint2 *>(&sm.l_first[st * 128u + 4u * lane]) = make_uint2(f0 | (f1 << 16), f2 | (f3 << 16));
            total += __shfl_sync(FULL_MASK, incl, 31);
        }
    }
    if (lane == 0) sm.l_first[n_ch] = (uint16_t)total;
    if (rev && total < need_c) return false;  // "MM tag refers to bases beyond sequence length": reported by the streaming path
    __syncwarp();

    // ---- listed bases, 32 per step, in list order (descending SEQ offsets on reversed alignments) ----
    const uint32_t pat = (rev ? 4u : 2u) * 0x11111111u;
    const int cg = rev ? -1 : 0;
    const uint32_t qs = R.pos, i_ref = qs - clip;
    uint32_t n_out = 0, last_pos = 0, n_mods = 0, n_behind_clip = 0;
    bool have_last = false, unsorted = false, implicit = false;
    // Two rows of 32 listed bases per step: the rows are resolved with straight-line, predicated code (two independent
    // chains of dependent shared-memory loads in flight per lane), then stored row by row, in list order.
    const uint32_t last_chunk_off = (n_bytes - 1u) & ~15u;
    const uint32_t want_nb = rev ? 2u : 4u;
    for (uint32_t k0 = 0; k0 < n_targets; k0 += 64) {
        bool r_found[2], r_ok[2], r_emit[2];
        uint32_t r_pos[2], r_cat[2], r_p[2];
#pragma unroll
        for (int row = 0; row < 2; row++) {
            const uint32_t k = k0 + (uint32_t)row * 32u + lane;
            const uint32_t kk = k < n_targets ? k : n_targets - 1u;
            const uint32_t rr = rank[kk];
            const bool found = k < n_targets && rr < total;  // ranks ascend: the bases that lie beyond SEQ form the tail of the list
            const uint32_t r = found ? rr : 0u;
            // chunk (scan order) that holds rank r: last c with l_first[c] <= r; unused entries are 0xffff
            // (the cursor is a pointer so that every step is load-with-immediate-offset, compare, predicated add)
            // (the range shrinks by its probed half whether or not the step is taken — len - len/2 is at least the half
            //  that stays — so the probe offsets are compile-time constants for any table size)
            const uint16_t *fq = sm.l_first;
#pragma unroll
            for (uint32_t ln = LEAN_CH; ln > 1; ln -= ln >> 1)
                if (fq[ln >> 1] <= r) fq += ln >> 1;
            const uint32_t c = (uint32_t)(fq - sm.l_first);
            const uint32_t f0 = fq[0], cnt = fq[1] - f0;
            const uint32_t chunk = rev ? n_ch - 1u - c : c;
            uint32_t n = rev ? cnt - 1u - (r - f0) : r - f0;  // index in base order inside the chunk
            const uint32_t coff = chunk * 16u < last_chunk_off ? chunk * 16u : last_chunk_off;
            const uint4 v = *reinterpret_cast<const uint4 *>(seq + coff);
            const uint32_t g0 = nib_eq_flags(v.x, pat), g1 = nib_eq_flags(v.y, pat), g2 = nib_eq_flags(v.z, pat),
                           g3 = nib_eq_flags(v.w, pat);
            const uint32_t c0 = (uint32_t)__popc(g0), c1 = c0 + (uint32_t)__popc(g1), c2 = c1 + (uint32_t)__popc(g2);
            const uint32_t wj = (n >= c0 ? 1u : 0u) + (n >= c1 ? 1u : 0u) + (n >= c2 ? 1u : 0u);
            const uint32_t f = wj == 0u ? g0 : (wj == 1u ? g1 : (wj == 2u ? g2 : g3));
            n -= wj == 0u ? 0u : (wj == 1u ? c0 : (wj == 2u ? c1 : c2));
            const uint32_t p = chunk * 32u + wj * 8u + select_base_in_word(f, n);
            // blockjoin.c:846-858: C must be followed by G; on reversed alignments SEQ shows the G, preceded by C
            const bool inner = found && p > 0 && p < len - 1;
            // the neighbouring base: inside the 16-byte chunk at hand (nibble i of a word holds base i ^ 1) unless the base
            // is the chunk's first / last one
            const uint32_t b = (p & 31u) + (rev ? 0xffffffffu : 1u);  // neighbour's index inside the chunk: -1 .. 32
            uint32_t nb;
            if (b < 32u) {
                const uint32_t w = b < 16u ? (b < 8u ? v.x : v.y) : (b < 24u ? v.z : v.w);
                nb = (w >> ((((b & 7u) ^ 1u)) << 2)) & 0xfu;
            } else nb = inner ? seq_nib(seq, rev ? p - 1u : p + 1u) : 0u;
            const bool ok = inner && nb == want_nb;
            if (inner && !ok) implicit = true;
            const uint32_t q = has_ml ? ml[ml_base + kk * stride + m_idx] : 255u;
            r_cat[row] = q < P.lo ? 1u : (q >= P.hi ? 0u : 2u);  // blockjoin.c:876-878
            // the operation whose inclusive trigger loop (blockjoin.c:663-665) handles the base: first one ending at or behind it
            const uint16_t *eq = &sm.l_end[0] - 1;  // (eq[h] is l_end[j + h - 1]; eq + 1 - l_end counts the entries < p)
#pragma unroll
            for (uint32_t ln = LEAN_MI; ln > 1; ln -= ln >> 1)
                if (eq[ln >> 1] < p) eq += ln >> 1;
            const uint32_t j = (uint32_t)(eq + 1 - sm.l_end);
            const int32_t o = sm.l_off[j < (uint32_t)LEAN_MI ? j : 0u];
            const bool at_clip = j0 && p == clip;  // a base right at the clip edge (blockjoin.c:640-652)
            const bool mapped = (p > clip || !j0) && j < n_mi && o != (int32_t)LEAN_DROP;
            r_found[row] = found;
            r_ok[row] = ok;
            r_p[row] = p;
            r_emit[row] = ok && (at_clip || mapped);
            r_pos[row] = at_clip ? qs + (uint32_t)cg : i_ref + p + (uint32_t)cg + (uint32_t)o;
        }
        bool done = false;
#pragma unroll
        for (int row = 0; row < 2; row++) {
            const bool emit = r_emit[row];
            const uint32_t pos = r_pos[row], cat = r_cat[row];
            const unsigned okm = __ballot_sync(FULL_MASK, r_ok[row]);
            n_mods += (uint32_t)__popc(okm);
            n_behind_clip += (uint32_t)__popc(__ballot_sync(FULL_MASK, r_ok[row] && r_p[row] > clip));
            // ordered, de-duplicated store (blockjoin.c:704-709: of the bases that land on one position the one latest in
            // the read sets the category).  Forward: ascending from slot 0; reversed: descending from the top slot.
            const unsigned em = __ballot_sync(FULL_MASK, emit);
            const unsigned below = em & ((1u << lane) - 1u);
            const int prev_lane = below ? 31 - __clz((int)below) : -1;
            uint32_t prev_pos = __shfl_sync(FULL_MASK, pos, prev_lane < 0 ? 0 : prev_lane);
            const bool has_prev = prev_lane >= 0 || have_last;
            if (prev_lane < 0) prev_pos = last_pos;
            const bool head = emit && !(has_prev && prev_pos == pos);
            if (head && has_prev && (rev ? pos > prev_pos : pos < prev_pos)) unsorted = true;
            const unsigned hm = __ballot_sync(FULL_MASK, head);
            const uint32_t run_idx = n_out + (uint32_t)__popc(hm & ((2u << lane) - 1u)) - 1u;  // a continuation of the carried run: n_out - 1
            if (!rev) {
                const unsigned above = em & ~((2u << lane) - 1u);
                const int next_lane = above ? __ffs((int)above) - 1 : -1;
                const uint32_t next_pos = __shfl_sync(FULL_MASK, pos, next_lane < 0 ? 0 : next_lane);
                const bool last_of_run = emit && (next_lane < 0 || next_pos != pos);
                if (head) opos[run_idx] = pos;
                if (last_of_run) ocat[run_idx] = (uint8_t)cat;
            } else if (head) {
                opos[cap - 1u - run_idx] = pos;
                ocat[cap - 1u - run_idx] = (uint8_t)cat;
            }
            n_out += (uint32_t)__popc(hm);
            if (em) {
                last_pos = __shfl_sync(FULL_MASK, pos, 31 - __clz((int)em));
                have_last = true;
            }
            if (__ballot_sync(FULL_MASK, r_found[row]) != FULL_MASK) done = true;
        }
        if (done) break;
    }
    const bool has_implicit = __any_sync(FULL_MASK, implicit);
    if (n_mods == 0) { *status_out = RS_LEAN | (has_implicit ? RS_HAS_IMPLICIT : 0u); *n_calls_out = 0; return true; }  // get_mod_poss_on_ref returns 0: record dropped
    if (j0 && n_behind_clip == 0) return false;                     // every listed base inside the clip: lingering-trigger quirk, streaming path
    __syncwarp();
    if (has_implicit) {
        // ---- implicit canonical calls (blockjoin.c:666-700, 727-761): a listed cytosine outside a CpG makes
        // get_mod_poss_on_ref fill in every unlisted CpG of the aligned stretches as unmethylated.  The walk over
        // CIGAR operations and kept mods stays sequential (executed by the whole warp, uniformly), as in the reference;
        // the CpG scan of the stretch between two mods is done by the 32 lanes.  The explicit calls stored above are
        // overwritten: the merged sequence is rebuilt from the kept mods.
        // the kept mods themselves (read offset, category), in list order — what get_mod_poss_on_ref is handed: the listed
        // bases are resolved once more, one row at a time (only records with implicit calls pay for this)
        {
            uint32_t nt = 0;
            for (uint32_t k0 = 0; k0 < n_targets; k0 += 32) {
                const uint32_t k = k0 + lane;
                const uint32_t kk = k < n_targets ? k : n_targets - 1u;
                const uint32_t rr = rank[kk];
                const bool found = k < n_targets && rr < total;  // ranks ascend: the bases that lie beyond SEQ form the tail of the list
                const uint32_t r = found ? rr : 0u;
                // chunk (scan order) that holds rank r: last c with l_first[c] <= r; unused entries are 0xffff
                // (the cursor is a pointer so that every step is load-with-immediate-offset, compare, predicated add)
                // (the range shrinks by its probed half whether or not the step is taken — len - len/2 is at least the half
                //  that stays — so the probe offsets are compile-time constants for any table size)
                const uint16_t *fq = sm.l_first;
#pragma unroll
                for (uint32_t ln = LEAN_CH; ln > 1; ln -= ln >> 1)
                    if (fq[ln >> 1] <= r) fq += ln >> 1;
                const uint32_t c = (uint32_t)(fq - sm.l_first);
                const uint32_t f0 = fq[0], cnt = fq[1] - f0;
                const uint32_t chunk = rev ? n_ch - 1u - c : c;
                uint32_t n = rev ? cnt - 1u - (r - f0) : r - f0;  // index in base order inside the chunk
                const uint32_t coff = chunk * 16u < last_chunk_off ? chunk * 16u : last_chunk_off;
                const uint4 v = *reinterpret_cast<const uint4 *>(seq + coff);
                const uint32_t g0 = nib_eq_flags(v.x, pat), g1 = nib_eq_flags(v.y, pat), g2 = nib_eq_flags(v.z, pat),
                               g3 = nib_eq_flags(v.w, pat);
                const uint32_t c0 = (uint32_t)__popc(g0), c1 = c0 + (uint32_t)__popc(g1), c2 = c1 + (uint32_t)__popc(g2);
                const uint32_t wj = (n >= c0 ? 1u : 0u) + (n >= c1 ? 1u : 0u) + (n >= c2 ? 1u : 0u);
                const uint32_t f = wj == 0u ? g0 : (wj == 1u ? g1 : (wj == 2u ? g2 : g3));
                n -= wj == 0u ? 0u : (wj == 1u ? c0 : (wj == 2u ? c1 : c2));
                const uint32_t p = chunk * 32u + wj * 8u + select_base_in_word(f, n);
                // blockjoin.c:846-858: C must be followed by G; on reversed alignments SEQ shows the G, preceded by C
                const bool inner = found && p > 0 && p < len - 1;
                // the neighbouring base: inside the 16-byte chunk at hand (nibble i of a word holds base i ^ 1) unless the base
                // is the chunk's first / last one
                const uint32_t b = (p & 31u) + (rev ? 0xffffffffu : 1u);  // neighbour's index inside the chunk: -1 .. 32
                uint32_t nb;
                if (b < 32u) {
                    const uint32_t w = b < 16u ? (b < 8u ? v.x : v.y) : (b < 24u ? v.z : v.w);
                    nb = (w >> ((((b & 7u) ^ 1u)) << 2)) & 0xfu;
                } else nb = inner ? seq_nib(seq, rev ? p - 1u : p + 1u) : 0u;
                const bool ok = inner && nb == want_nb;
                const uint32_t q = has_ml ? ml[ml_base + kk * stride + m_idx] : 255u;
                const uint32_t cat = q < P.lo ? 1u : (q >= P.hi ? 0u : 2u);
                const unsigned okm = __ballot_sync(FULL_MASK, ok);
                if (ok) {
                    const uint32_t ti = nt + (uint32_t)__popc(okm & ((1u << lane) - 1u));
                    trig_p[ti] = p;
                    trig_cat[ti] = (uint8_t)cat;
                }
                nt += (uint32_t)__popc(okm);
                if (__ballot_sync(FULL_MASK, found) != FULL_MASK) break;
            }
            __syncwarp();
        }
        uint32_t n = 0;
        const uint32_t wst = implicit_walk(cigar, n_cigar, seq, len, qs, rev, trig_p, trig_cat, n_mods, rev, opos, ocat, cap, &n);
        *n_calls_out = n;
        *status_out = RS_LEAN | RS_HAS_IMPLICIT | wst;
        return true;
    }
    if (rev && n_out && cap - n_out) {  // slide the calls down to the front of the record's slots
        const uint32_t shift = cap - n_out;
        for (uint32_t i0 = 0; i0 < n_out; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t a = 0;
            uint8_t b = 0;
            if (i < n_out) { a = opos[i + shift]; b = ocat[i + shift]; }
            __syncwarp();
            if (i < n_out) { opos[i] = a; ocat[i] = b; }
            __syncwarp();
        }
    }
    *n_calls_out = n_out;
    *status_out = RS_KEPT | RS_LEAN | (__any_sync(FULL_MASK, unsorted) ? RS_UNSORTED : 0u);
    return true;
}

// ---------------------------------------------------------------------------------------------
// General sequential path (lane 0): any number of C+m streams, more than N_MODS streams, implicit
// canonical calls.  Follows the reference statement by statement.
// ---------------------------------------------------------------------------------------------
constexpr int GEN_MAXSEG = 64;

struct GenSeg {
    uint32_t list_begin, list_end, n_delta, total, ml_base;
    uint32_t n_codes;
    uint32_t code_begin;   // offset of the first code character
    uint8_t canon, is_chebi;
    // walk state
    uint32_t cursor;       // forward: offset of the next ','; reverse: offset just past the current delta
    uint32_t next_target;  // index among matching bases (from the left of SEQ) of the next listed base
    uint32_t match_idx;
    uint32_t k;            // forward: next delta index; reverse: deltas still to hand out
};

__device__ uint32_t gen_parse_uint(const uint8_t *s, uint32_t b, uint32_t e, uint32_t *next) {
    uint32_t v = 0;
    while (b < e && is_digit(s[b])) { v = v >= DEC_SAT / 10 ? DEC_SAT : v * 10 + (s[b] - '0'); b++; }
    *next = b;
    return v >= DEC_SAT ? DEC_SAT - 1 : v;
}

struct CallSink {
    uint32_t *pos;
    uint8_t *cat;
    uint32_t n, cap;
    bool unsorted;
    __device__ void push(uint32_t p, uint8_t c) {
        if (n > 0 && n <= cap && p <= pos[n - 1]) unsorted = true;
        if (n < cap) { pos[n] = p; cat[n] = c; }
        n++;
    }
    __device__ bool last_is(uint32_t p) const { return n > 0 && n <= cap && pos[n - 1] == p; }
    __device__ void set_last_cat(uint8_t c) { if (n > 0 && n <= cap) cat[n - 1] = c; }
};

__device__ void gen_implicit_fill(CallSink &out, const uint8_t *seq, uint32_t len, uint32_t from, uint32_t until,
                                  uint32_t i_ref, int32_t offset) {
    for (uint32_t t = from; t < until; t++) {
        if (t < len - 1 && seq_nib(seq, t) == 2u && seq_nib(seq, t + 1) == 4u) {
            uint32_t p = i_ref + t + (uint32_t)offset;
            if (!out.last_is(p)) out.push(p, 1);
            t++;
        }
    }
}

__device__ uint32_t decode_generic(const DecodeParams &P, const ReadRec &R, GenSeg *segs, uint32_t *n_calls_out) {
    const uint8_t *blob = P.blob;
    const uint32_t *cigar = reinterpret_cast<const uint32_t *>(blob + (size_t)R.cigar_off * 16);
    const uint8_t *seq = blob + (size_t)R.seq_off * 16;
    const uint8_t *mm = blob + (size_t)R.mm_off * 16;
    const uint8_t *ml = blob + (size_t)R.ml_off * 16;
    const uint32_t len = R.l_qseq, mm_len = R.mm_len;
    const bool rev = (R.flags & 16u) != 0, has_ml = (R.flags & RF_HAS_ML) != 0;
    uint32_t *mpos = P.tmp_mpos + R.calls_off;
    uint8_t *mcat = P.tmp_mcat + R.calls_off;
    const uint32_t cap = R.calls_cap;
    uint32_t status = RS_SLOWPATH;
    *n_calls_out = 0;
    int n_seg = 0;
    bool bad = false;
    if (!(R.flags & RF_HAS_MM)) n_seg = 0;
    else if (R.flags & RF_MALFORMED) bad = true;
    else {
        uint32_t p = 0, ml_used = 0, n_streams = 0;
        while (p < mm_len && !bad) {
            if (n_seg >= GEN_MAXSEG) { *n_calls_out = 0; return status | RS_MM_ERROR | RS_OVERFLOW | RS_FATAL_CIGAR; }
            GenSeg &g = segs[n_seg];
            int canon = base_code_of(mm[p]);
            if (canon < 0) { bad = true; break; }
            p++;
            if (p >= mm_len || (mm[p] != '+' && mm[p] != '-')) { bad = true; break; }
            p++;
            g.code_begin = p;
            g.is_chebi = 0;
            g.n_codes = 0;
            if (p < mm_len && is_digit(mm[p])) { while (p < mm_len && is_digit(mm[p])) p++; g.n_codes = 1; g.is_chebi = 1; }
            else {
                while (p < mm_len && is_alpha(mm[p])) { g.n_codes++; p++; }
                if (p >= mm_len) { bad = true; break; }
            }
            if (p < mm_len && (mm[p] == '.' || mm[p] == '?')) p++;
            else if (p >= mm_len || (mm[p] != ',' && mm[p] != ';')) { bad = true; break; }
            if (g.n_codes > 0 && n_streams + g.n_codes >= 256) { bad = true; break; }
            g.list_begin = p;
            g.n_delta = 0;
            g.total = 0;
            while (p < mm_len && mm[p] == ',') {
                p++;
                if (p >= mm_len || !is_digit(mm[p])) { bad = true; break; }
                uint32_t nx;
                uint32_t v = gen_parse_uint(mm, p, mm_len, &nx);
                p = nx;
                g.n_delta++;
                g.total = sat_add(g.total, v + 1);
            }
            if (bad) break;
            if (p >= mm_len || mm[p] != ';') { bad = true; break; }
            g.list_end = p;
            p++;
            g.canon = (uint8_t)canon;
            g.ml_base = ml_used;
            if (has_ml && ml_used + g.n_delta * g.n_codes > R.ml_len) { bad = true; break; }
            ml_used += g.n_delta * g.n_codes;
            n_streams += g.n_codes;
            n_seg++;
        }
        if (!bad && has_ml && ml_used != R.ml_len) bad = true;
        if (!bad && rev) {
            uint32_t freq[16];
            for (int i = 0; i < 16; i++) freq[i] = 0;
            for (uint32_t i = 0; i < len; i++) freq[seq_nib(seq, i)]++;
            for (int s = 0; s < n_seg; s++) {
                GenSeg &g = segs[s];
                if (g.n_codes == 0) continue;
                uint32_t f = freq[comp_code(g.canon)];
                if (g.total > f) { bad = true; break; }
                g.next_target = f - g.total;  // lead
            }
        }
    }
    if (bad) { status |= RS_MM_ERROR; n_seg = 0; }
    // ---- walk SEQ, blockjoin.c:832-882 on top of the stateful iterator ----
    for (int s = 0; s < n_seg; s++) {
        GenSeg &g = segs[s];
        g.match_idx = 0;
        if (!rev) {
            g.k = 0;
            g.cursor = g.list_begin;
            if (g.n_delta > 0) {
                uint32_t nx;
                uint32_t d = gen_parse_uint(mm, g.cursor + 1, mm_len, &nx);
                g.cursor = nx;
                g.next_target = d;
            }
        } else {
            g.k = g.n_delta;          // deltas not yet handed out; the left-most listed base is delta n-1
            g.cursor = g.list_end;    // just past the last delta
        }
    }
    uint32_t n_mods = 0;
    bool has_implicit = false, mod_overflow = false;
    bool any_left = false;
    for (int s = 0; s < n_seg; s++) if (segs[s].n_delta > 0 && segs[s].n_codes > 0) any_left = true;
    for (uint32_t p = 0; p < len && any_left; p++) {
        uint32_t code = seq_nib(seq, p);
        uint32_t mcode = rev ? comp_code(code) : code;
        uint32_t n_here = 0;
        uint32_t first_new = n_mods;
        bool imp_here = false;
        for (int s = 0; s < n_seg; s++) {
            GenSeg &g = segs[s];
            if (g.n_codes == 0 || g.n_delta == 0) continue;
            if (mcode != g.canon && g.canon != 15) continue;
            bool done = !rev ? g.k >= g.n_delta : g.k == 0;
            if (!done && g.match_idx == g.next_target) {
                uint32_t which = !rev ? g.k : g.k - 1;
                n_here += g.n_codes;
                if (g.canon == 2 && !g.is_chebi && p < len - 1 && p > 0) {
                    for (uint32_t c = 0; c < g.n_codes; c++) {
                        if (mm[g.code_begin + c] != 'm') continue;
                        bool ok = code == 2u ? seq_nib(seq, p + 1) == 4u : seq_nib(seq, p - 1) == 2u;
                        if (!ok) { imp_here = true; continue; }
                        uint32_t q = has_ml ? ml[g.ml_base + which * g.n_codes + c] : 255u;
                        if (n_mods < cap) { mpos[n_mods] = p; mcat[n_mods] = q < P.lo ? 1 : (q >= P.hi ? 0 : 2); }
                        else mod_overflow = true;
                        n_mods++;
                    }
                }
                // advance to the next listed base of this segment
                if (!rev) {
                    g.k++;
                    if (g.k < g.n_delta) {
                        uint32_t nx;
                        uint32_t d = gen_parse_uint(mm, g.cursor + 1, mm_len, &nx);
                        g.cursor = nx;
                        g.next_target = sat_add(g.match_idx + 1, d);
                    }
                } else {
                    // skip count before the next (further right) listed base is the delta we just consumed
                    uint32_t b = g.cursor;  // just past delta `which`
                    while (b > g.list_begin && mm[b - 1] != ',') b--;
                    uint32_t nx;
                    uint32_t d = gen_parse_uint(mm, b, mm_len, &nx);
                    g.cursor = b - 1;       // the ',' in front of it = just past delta which-1
                    g.k--;
                    g.next_target = sat_add(g.match_idx + 1, d);
                }
            }
            g.match_idx++;
        }
        if (n_here > N_MODS_LIMIT) n_mods = first_new;  // "mod reading buffer was length 10 but iter saw n": position skipped
        else if (imp_here) has_implicit = true;
    }
    if (has_implicit) status |= RS_HAS_IMPLICIT;
    if (mod_overflow) { *n_calls_out = n_mods; return status | RS_OVERFLOW; }

    // ---- get_mod_poss_on_ref, blockjoin.c:605-792 ----
    const uint32_t n_cigar = R.n_cigar;
    if (n_cigar == 0 || n_mods == 0) return status;
    CallSink out;
    out.pos = P.calls_pos + R.calls_off;
    out.cat = P.calls_cat + R.calls_off;
    out.n = 0; out.cap = cap; out.unsorted = false;
    const int cg = rev ? -1 : 0;
    uint32_t i_read = 0, i_ref = R.pos, it = 0, next = mpos[0];
    uint8_t nq = mcat[0];
    uint32_t ic = 0;
    if ((cigar[0] & 15u) == 4u) {
        i_read = cigar[0] >> 4;
        while (next < i_read) {
            it++;
            if (it < n_mods) { next = mpos[it]; nq = mcat[it]; } else break;
        }
        if (next == i_read) {
            out.push(i_ref + (uint32_t)cg, nq);
            it++;
            if (it < n_mods) { next = mpos[it]; nq = mcat[it]; }
        }
        i_ref -= cigar[0] >> 4;
        ic = 1;
    }
    int32_t offset = 0;
    bool fatal = false;
    for (; ic < n_cigar; ic++) {
        uint32_t op = cigar[ic] & 15u, L = cigar[ic] >> 4;
        if (op <= 1u) {
            uint32_t pos_canonical = i_read;
            while (i_read + L >= next) {
                if (op == 0u && next != 0xffffffffu) {
                    if (has_implicit) {
                        uint32_t until = next - 1 < i_read + L ? next - 1 : i_read + L;
                        gen_implicit_fill(out, seq, len, pos_canonical, until, i_ref, offset);
                    }
                    uint32_t pt = i_ref + next + (uint32_t)cg + (uint32_t)offset;
                    if (out.last_is(pt)) out.set_last_cat(nq); else out.push(pt, nq);
                    pos_canonical = cg == 0 ? next + 1 : next + 2;
                }
                it++;
                if (it >= n_mods) { next = 0xffffffffu; break; }
                next = mpos[it];
                nq = mcat[it];
            }
            if (op == 0u) {
                if (has_implicit) gen_implicit_fill(out, seq, len, pos_canonical, i_read + L, i_ref, offset);
                i_read += L;
            } else { i_read += L; offset -= (int32_t)L; }
        } else if (op == 2u) offset += (int32_t)L;
        else if (op == 3u || op == 4u) break;
        else { fatal = true; break; }
    }
    if (fatal) return status | RS_FATAL_CIGAR;
    *n_calls_out = out.n;
    if (out.n > cap) status |= RS_OVERFLOW;
    if (out.unsorted) status |= RS_UNSORTED;
    return status | RS_KEPT;
}

// ---------------------------------------------------------------------------------------------
// Kernel: one warp per record.
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(DecodeWarpSmem) >= sizeof(GenSeg) * GEN_MAXSEG, "the general path reuses the warp's staging area");

// Persistent warps: every warp takes the next record off a queue (longest records first), so neither the
// spread of record lengths inside a CTA nor the last wave of the grid leaves warp slots idle.
__global__ void __launch_bounds__(DEC_WARPS * 32, POMFRET_DEC_MIN_CTAS) decode_kernel(DecodeParams P) {
    __shared__ DecodeWarpSmem smem[DEC_WARPS];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    DecodeWarpSmem &sm = smem[warp];
    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(P.next, 1u);
        qi = __shfl_sync(FULL_MASK, qi, 0);
        if (qi >= P.n_queue) return;  // whole warp leaves together
        const uint32_t ri = P.order ? P.order[qi] : qi;
        const ReadRec &R = P.reads[ri];  // read-only for the whole launch: fields are fetched where they are used
        uint32_t n_calls = 0, status = 0, rlen = 0;
        if (!P.no_lean && decode_lean(P, R, sm, &status, &n_calls, &rlen)) {
            if (R.flags & 4u) rlen = 0;
            if (rlen == 0) rlen = 1;
        } else {
            __syncwarp();
            // reference span of the alignment (bam_endpos): M, D, N, =, X consume the reference
            const uint32_t *cigar = reinterpret_cast<const uint32_t *>(P.blob + (size_t)R.cigar_off * 16);
            rlen = 0;
            for (uint32_t i = lane; i < R.n_cigar; i += 32) {
                uint32_t c = cigar[i], op = c & 15u;
                if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) rlen += c >> 4;
            }
            rlen = warp_sum(rlen);
            if (R.flags & 4u) rlen = 0;
            if (rlen == 0) rlen = 1;
            bool need_generic = false;
            status = decode_fast(P, R, sm, &n_calls, &need_generic);
            need_generic = __any_sync(FULL_MASK, need_generic);
            if (need_generic) {
                __syncwarp();
                if (P.generic_list) {
                    // The general path is sequential: instead of idling 31 lanes on it, the record goes on a list that
                    // decode_generic_kernel works through with one record per THREAD (32 records per warp).
                    if (lane == 0) {
                        P.generic_list[atomicAdd(P.n_generic, 1u)] = ri;
                        P.r_end[ri] = R.pos + rlen;
                    }
                    __syncwarp();
                    continue;
                }
                if (lane == 0) status = decode_generic(P, R, reinterpret_cast<GenSeg *>(&sm), &n_calls);
                status = __shfl_sync(FULL_MASK, status, 0);
                n_calls = __shfl_sync(FULL_MASK, n_calls, 0);
            }
        }
        if (lane == 0) {
            if (status & RS_OVERFLOW) atomicAdd(P.n_overflow, 1u);
            P.r_ncalls[ri] = (status & RS_KEPT) || (status & RS_OVERFLOW) ? n_calls : 0;
            P.r_status[ri] = status;
            P.r_end[ri] = R.pos + rlen;
        }
        __syncwarp();
    }
}

// The records decode_kernel put aside for the general sequential path (several C+m streams, more than N_MODS streams,
// implicit canonical calls, blockjoin.c:666-700): one record per thread, segment tables in local memory.  An all-context
// 5mC data set sends every record here; one record per warp (lane 0) left 31 of 32 lanes of every warp slot idle.
constexpr int GEN_THREADS = 64;
__global__ void __launch_bounds__(GEN_THREADS) decode_generic_kernel(DecodeParams P) {
    const uint32_t i = blockIdx.x * GEN_THREADS + threadIdx.x;
    if (i >= *P.n_generic) return;
    const uint32_t ri = P.generic_list[i];
    const ReadRec &R = P.reads[ri];
    GenSeg segs[GEN_MAXSEG];
    uint32_t n_calls = 0;
    const uint32_t status = decode_generic(P, R, segs, &n_calls);
    if (status & RS_OVERFLOW) atomicAdd(P.n_overflow, 1u);
    P.r_ncalls[ri] = (status & RS_KEPT) || (status & RS_OVERFLOW) ? n_calls : 0;
    P.r_status[ri] = status;
}

// A record that lies in two windows occupies two slots of the batch but is decoded once: the later slot takes
// the result (call count, status, reference end) of the earlier one; their ReadRec already share the call slots.
__global__ void share_decoded_kernel(const uint32_t *dup_of, uint32_t n, uint32_t *r_ncalls, uint32_t *r_status, uint32_t *r_end) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = dup_of[i];
    if (j == 0xffffffffu) return;
    r_ncalls[i] = r_ncalls[j]; r_status[i] = r_status[j]; r_end[i] = r_end[j];
}

}  // namespace pomfret_gpu
#endif
