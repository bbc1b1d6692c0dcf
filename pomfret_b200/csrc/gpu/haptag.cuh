// (c) Read haplotagging against the phased VCF for untagged BAMs (`-u` / --bam-is-untagged).
// One warp per alignment record.
//
// Replaces parse_variants_for_one_read (reference blockjoin.c:1545-1691) and
// haptag_one_read_with_variants (blockjoin.c:1693-1840); the known-variant set is what
// insert_variant_from_vcf_line builds (blockjoin.c:1432-1543), the per-read i_left cursor
// (blockjoin.c:1716-1720) is computed by the caller, first-alignment-wins (1880-1889) stays on the host.
//
//   A. CIGAR scan (32 ops per step, warp prefix sums): reference span, insertion list with, per
//      insertion, its reference position, length, read position and read position net of earlier
//      insertions (that is what the MD walk's "skip insertions" rule compares against).
//   B. MD scan (512 chars per step): every lane classifies 16 characters; the tokenizer state
//      (inside a number / inside a ^deletion run / neither) follows from a warp scan of per-lane summaries, then each lane
//      emits its tokens (mismatch with the read base, or deletion with the reference letters) with
//      reference / read offsets from warp prefix sums.
//   C. Vote walk, one known variant per lane: the reference merges known and read variants with a 64-bit radix
//      sort and walks the buffer; both read-variant lists are already position sorted, so the walk visits only the
//      known variants and finds its neighbours in the merged order by binary search.
#ifndef POMFRET_GPU_HAPTAG_CUH
#define POMFRET_GPU_HAPTAG_CUH
#include "gpu_rt.h"
#include "types.h"
#include "decode.cuh"

namespace pomfret_gpu {

constexpr int HAP_WARPS = 4;
constexpr int HAP_MAX_UNSORTED = 64;

struct KnownVar {   // mirrors pomfret_gpu_variant
    uint32_t pos, len;
    uint8_t op, haptag;
    uint16_t reserved;
    uint32_t bases_off;
};

struct HaptagParams {
    const ReadRec *reads;
    uint32_t n_reads;
    const uint8_t *blob;
    const KnownVar *known;
    uint32_t n_known;
    const uint8_t *bases;
    const uint32_t *known_first;
    // scratch, sliced per read by scr_off (insertions: n_cigar slots, MD variants: md_len slots)
    const uint32_t *ins_off, *mdv_off;
    uint32_t *ins_ref, *ins_len, *ins_self, *ins_q, *ins_cum;
    uint32_t *mv_pos, *mv_info, *mv_aux;
    uint8_t *out_tag;
    int32_t *out_status;
    int32_t *out_votes;  // 2 per read
    uint32_t *out_counts; // 2 per read: insertions, MD variants left in the scratch lists (nullptr: not wanted)
};

struct HapWarpSmem {
    __align__(16) uint8_t buf[32 + 512 + 16];
};

__device__ __forceinline__ uint8_t nt4_of_nib(uint32_t c) { return c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4; }
__device__ __forceinline__ uint8_t nt4_of_char(uint32_t ch) {
    switch (ch) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3; default: return 4;
    }
}
// md_op_table, blockjoin.c:94-115
__device__ __forceinline__ int md_class(uint32_t ch) {
    if (ch >= '0' && ch <= '9') return 0;
    if (ch == '^') return 1;
    switch (ch) {
    case 'A': case 'C': case 'G': case 'T': case 'U': case 'N':
    case 'a': case 'c': case 'g': case 't': case 'u': case 'n': return 2;
    default: return 4;
    }
}

// tokenizer state handed from lane to lane
struct MdState {
    uint32_t mode;       // 0 none, 1 inside a number, 2 inside a ^ run
    uint32_t num;        // value of the number so far
    uint32_t del_start;  // offset of the '^'
    uint32_t since;      // mismatches emitted since the last number ended
};

// Summary of a stretch of MD characters, enough to tell the tokenizer state behind it:
//   the part behind the stretch's last digit (the whole stretch if it has none): offset of its first '^' and the
//   letters in front of that caret (all its letters if there is no caret) — a caret opens a deletion run that only a
//   digit ends, letters in front of it are mismatches counted "since the last number";
//   the digit run the stretch ends with: its value (32-bit wrap-around like the sequential num*10+d) and, for a
//   stretch of digits only, 10^length so that a run can continue across stretches.
constexpr uint32_t MDS_HAS_DIGIT = 1u, MDS_ENDS_DIGIT = 2u, MDS_KIND_DIGITS = 4u, MDS_KIND_MIXED = 8u;  // no kind bit: empty stretch
constexpr uint32_t MDS_NONE = 0xffffffffu;
struct MdSum {
    uint32_t flags, caret, letters, v, p10;
};

__device__ __forceinline__ MdSum md_combine(const MdSum &A, const MdSum &B) {  // A, then B
    MdSum C;
    uint32_t f;
    if (B.flags & MDS_HAS_DIGIT) { f = MDS_HAS_DIGIT; C.caret = B.caret; C.letters = B.letters; }
    else {
        f = A.flags & MDS_HAS_DIGIT;
        const bool a_caret = A.caret != MDS_NONE;
        C.caret = a_caret ? A.caret : B.caret;
        C.letters = a_caret ? A.letters : A.letters + B.letters;
    }
    const uint32_t bk = B.flags & (MDS_KIND_DIGITS | MDS_KIND_MIXED), ak = A.flags & (MDS_KIND_DIGITS | MDS_KIND_MIXED);
    if (bk == 0u) { f |= A.flags & (MDS_ENDS_DIGIT | MDS_KIND_DIGITS | MDS_KIND_MIXED); C.v = A.v; C.p10 = A.p10; }
    else if (bk == MDS_KIND_DIGITS && ak != 0u) {
        f |= MDS_ENDS_DIGIT | ak;
        C.v = (A.flags & MDS_ENDS_DIGIT) ? A.v * B.p10 + B.v : B.v;
        C.p10 = A.p10 * B.p10;
    } else { f |= B.flags & (MDS_ENDS_DIGIT | MDS_KIND_DIGITS | MDS_KIND_MIXED); C.v = B.v; C.p10 = B.p10; }
    C.flags = f;
    return C;
}

// classes of a lane's 16 characters, two bits each (md_class 0 digit, 1 caret, 2 letter; 3 stands for class 4, invalid)
__device__ __forceinline__ uint32_t md_class_word(const uint32_t (&w)[4], uint32_t nv) {
    uint32_t cw = 0;
#pragma unroll
    for (uint32_t i = 0; i < 16; i++) {
        const uint32_t c = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
        const uint32_t li = (c | 0x20u) - 'a';  // letters: case folded, index in the alphabet
        const bool letter = li < 26u && ((0x00182045u >> li) & 1u);  // a c g n t u
        const uint32_t cl = (c - '0') <= 9u ? 0u : (c == '^' ? 1u : (letter ? 2u : 3u));
        cw |= (i < nv ? cl : 0u) << (2 * i);
    }
    return cw;
}

__device__ __forceinline__ MdSum md_summarise(const uint32_t (&w)[4], uint32_t cw, uint32_t nv, uint32_t off) {
    MdSum S;
    S.flags = 0; S.caret = MDS_NONE; S.letters = 0; S.v = 0; S.p10 = 1;
    bool prev_digit = false, nondigit = false;
    for (uint32_t i = 0; i < nv; i++) {
        const uint32_t c = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
        const uint32_t cl = (cw >> (2 * i)) & 3u;
        if (cl == 0) {
            S.flags |= MDS_HAS_DIGIT; S.caret = MDS_NONE; S.letters = 0;
            if (prev_digit) { S.v = S.v * 10u + (c - '0'); S.p10 *= 10u; } else { S.v = c - '0'; S.p10 = 10u; }
            prev_digit = true;
        } else {
            prev_digit = false; nondigit = true;
            if (S.caret == MDS_NONE) { if (cl == 1) S.caret = off + i; else if (cl == 2) S.letters++; }
        }
    }
    if (nv) S.flags |= (prev_digit ? MDS_ENDS_DIGIT : 0u) | (nondigit ? MDS_KIND_MIXED : MDS_KIND_DIGITS);
    return S;
}

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(HAP_WARPS * 32) haptag_kernel(HaptagParams P) {
    __shared__ HapWarpSmem smem[HAP_WARPS];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t ri = blockIdx.x * HAP_WARPS + warp;
    if (ri >= P.n_reads) return;
    const ReadRec R = P.reads[ri];
    HapWarpSmem &sm = smem[warp];
    const uint32_t *cigar = reinterpret_cast<const uint32_t *>(P.blob + (size_t)R.cigar_off * 16);
    const uint8_t *seq = P.blob + (size_t)R.seq_off * 16;
    const uint8_t *md = P.blob + (size_t)R.md_off * 16;
    uint32_t *ins_ref = P.ins_ref + P.ins_off[ri], *ins_len = P.ins_len + P.ins_off[ri];
    uint32_t *ins_self = P.ins_self + P.ins_off[ri], *ins_q = P.ins_q + P.ins_off[ri], *ins_cum = P.ins_cum + P.ins_off[ri];
    uint32_t *mv_pos = P.mv_pos + P.mdv_off[ri], *mv_info = P.mv_info + P.mdv_off[ri], *mv_aux = P.mv_aux + P.mdv_off[ri];

    // ---------------- A. CIGAR ----------------
    uint32_t ref_run = R.pos, self_run = 0, ins_total = 0, n_ins = 0, rlen = 0;
    uint32_t self_start = 0;
    for (uint32_t base = 0; base < R.n_cigar; base += 32) {
        const uint32_t i = base + lane;
        uint32_t op = 15, L = 0;
        if (i < R.n_cigar) { uint32_t c = cigar[i]; op = c & 15u; L = c >> 4; }
        // pass 1 of the reference: N,D advance ref; S,I advance read; M,=,X advance both (blockjoin.c:1564-1589)
        const uint32_t dref = (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) ? L : 0u;
        const uint32_t dself = (op == 0u || op == 1u || op == 4u || op == 7u || op == 8u) ? L : 0u;
        const uint32_t dins = op == 1u ? L : 0u;
        uint32_t iref = warp_inclusive_sum(dref), iself = warp_inclusive_sum(dself), iins = warp_inclusive_sum(dins);
        unsigned im = __ballot_sync(FULL_MASK, op == 1u);
        if (op == 1u) {
            uint32_t k = n_ins + __popc(im & ((1u << lane) - 1u));
            uint32_t sp = self_run + iself - dself;
            uint32_t before = ins_total + iins - dins;
            ins_ref[k] = ref_run + iref - dref;
            ins_len[k] = L;
            ins_self[k] = sp;
            ins_q[k] = sp - before;
            ins_cum[k] = before;
        }
        if (i == 0 && op == 4u) self_start = L;
        n_ins += __popc(im);
        ref_run += __shfl_sync(FULL_MASK, iref, 31);
        self_run += __shfl_sync(FULL_MASK, iself, 31);
        ins_total += __shfl_sync(FULL_MASK, iins, 31);
    }
    self_start = __shfl_sync(FULL_MASK, self_start, 0);
    rlen = ref_run - R.pos;
    if (rlen == 0) rlen = 1;
    const uint32_t end_pos = R.pos + rlen;
    __syncwarp();

    // ---------------- B. MD ----------------
    int32_t status = 0;
    uint32_t n_mv = 0;
    if (!(R.flags & RF_HAS_MD)) status = -9;  // assert(tagd), blockjoin.c:1596
    else if (R.md_len > 0) {
        MdSum carry;
        carry.flags = 0; carry.caret = MDS_NONE; carry.letters = 0; carry.v = 0; carry.p10 = 1;
        uint32_t ref_base = R.pos, b_base = self_start;  // running reference / insertion-free read offsets
        bool bad = false;
        for (uint32_t gbase = 0; gbase < R.md_len; gbase += 512) {
            const uint32_t off = gbase + lane * 16;
            uint4 v = *reinterpret_cast<const uint4 *>(md + off);  // padded blob: always readable
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t nv = off >= R.md_len ? 0u : (R.md_len - off < 16u ? R.md_len - off : 16u);
            // The tokenizer state a lane starts in follows from the characters since the last digit in front of it
            // (after a digit the state is "inside a number", whatever came before): summarise every lane's 16
            // characters, combine the summaries with one warp scan (md_combine is associative), and read the
            // state off the exclusive prefix.
            const uint32_t cw = md_class_word(w, nv);
            MdSum mine = md_summarise(w, cw, nv, off);
            MdSum incl = mine;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                MdSum up;
                up.flags = __shfl_up_sync(FULL_MASK, incl.flags, d);
                up.caret = __shfl_up_sync(FULL_MASK, incl.caret, d);
                up.letters = __shfl_up_sync(FULL_MASK, incl.letters, d);
                up.v = __shfl_up_sync(FULL_MASK, incl.v, d);
                up.p10 = __shfl_up_sync(FULL_MASK, incl.p10, d);
                if ((int)lane >= d) incl = md_combine(up, incl);
            }
            MdSum excl;
            excl.flags = __shfl_up_sync(FULL_MASK, incl.flags, 1);
            excl.caret = __shfl_up_sync(FULL_MASK, incl.caret, 1);
            excl.letters = __shfl_up_sync(FULL_MASK, incl.letters, 1);
            excl.v = __shfl_up_sync(FULL_MASK, incl.v, 1);
            excl.p10 = __shfl_up_sync(FULL_MASK, incl.p10, 1);
            excl = lane == 0 ? carry : md_combine(carry, excl);
            MdSum last;
            last.flags = __shfl_sync(FULL_MASK, incl.flags, 31);
            last.caret = __shfl_sync(FULL_MASK, incl.caret, 31);
            last.letters = __shfl_sync(FULL_MASK, incl.letters, 31);
            last.v = __shfl_sync(FULL_MASK, incl.v, 31);
            last.p10 = __shfl_sync(FULL_MASK, incl.p10, 31);
            carry = md_combine(carry, last);
            MdState in;
            in.mode = 0; in.num = 0; in.del_start = 0; in.since = 0;
            if (excl.flags & MDS_ENDS_DIGIT) { in.mode = 1; in.num = excl.v; }
            else if (excl.caret != MDS_NONE) { in.mode = 2; in.del_start = excl.caret; }
            else in.since = excl.letters;
            // every lane now knows its incoming state `in`: count what it emits
            uint32_t n_tok = 0, ref_adv = 0, b_adv = 0;
            {
                MdState s = in;
                for (uint32_t i = 0; i < nv; i++) {
                    const uint32_t c = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
                    const uint32_t cl = (cw >> (2 * i)) & 3u;
                    if (cl == 3) bad = true;
                    if (cl == 0) {
                        if (s.mode == 2) { uint32_t dl = off + i - s.del_start - 1; n_tok++; ref_adv += dl; s.mode = 1; s.num = c - '0'; }
                        else if (s.mode == 1) s.num = s.num * 10 + (c - '0');
                        else { s.mode = 1; s.num = c - '0'; }
                    } else {
                        if (s.mode == 1) { ref_adv += s.num; b_adv += s.num; s.mode = 0; s.since = 0; }
                        if (s.mode == 2) {}
                        else if (cl == 1) { s.mode = 2; s.del_start = off + i; }
                        else if (cl == 2) { n_tok++; ref_adv++; b_adv++; s.since++; }
                    }
                }
            }
            uint32_t i_tok = warp_inclusive_sum(n_tok), i_ref = warp_inclusive_sum(ref_adv), i_b = warp_inclusive_sum(b_adv);
            // emit
            {
                MdState s = in;
                uint32_t slot = n_mv + i_tok - n_tok;
                uint32_t rp = ref_base + i_ref - ref_adv;
                uint32_t bp = b_base + i_b - b_adv;
                for (uint32_t i = 0; i < nv; i++) {
                    const uint32_t c = (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
                    const uint32_t cl = (cw >> (2 * i)) & 3u;
                    if (cl == 0) {
                        if (s.mode == 2) {
                            uint32_t dl = off + i - s.del_start - 1;
                            mv_pos[slot] = rp; mv_info[slot] = (dl << 1) | 1u; mv_aux[slot] = s.del_start + 1;
                            slot++; rp += dl;
                            s.mode = 1; s.num = c - '0';
                        } else if (s.mode == 1) s.num = s.num * 10 + (c - '0');
                        else { s.mode = 1; s.num = c - '0'; }
                    } else {
                        if (s.mode == 1) { rp += s.num; bp += s.num; s.mode = 0; s.since = 0; }
                        if (s.mode == 2) {}
                        else if (cl == 1) { s.mode = 2; s.del_start = off + i; }
                        else if (cl == 2) {
                            // (the read base is looked up below, one mismatch per lane: parked here are B and the
                            //  mismatches since the last number ended)
                            mv_pos[slot] = rp; mv_info[slot] = (s.since << 2) | (1u << 1); mv_aux[slot] = bp;
                            slot++; rp++; bp++; s.since++;
                        }
                    }
                }
            }
            // read base of every mismatch of this step, one per lane: self_pos = B + (insertions whose net position
            // lies before the B reached when the last number ended), blockjoin.c:1628-1635, 1652
            __syncwarp();
            {
                const uint32_t n_new = __shfl_sync(FULL_MASK, i_tok, 31);
                for (uint32_t t0 = 0; t0 < n_new; t0 += 32) {
                    const uint32_t t = t0 + lane;
                    if (t >= n_new) continue;
                    const uint32_t slot = n_mv + t, info = mv_info[slot];
                    if (info & 1u) continue;  // a deletion
                    const uint32_t bp = mv_aux[slot], b_gap = bp - (info >> 2);
                    const uint32_t K = lower_bound_u32(ins_q, n_ins, b_gap);
                    const uint32_t skipped = K == 0 ? 0 : ins_cum[K - 1] + ins_len[K - 1];
                    const uint32_t sp = bp + skipped;
                    mv_info[slot] = 1u << 1;
                    mv_aux[slot] = sp < R.l_qseq ? nt4_of_nib(seq_nib(seq, sp)) : 4u;
                }
            }
            __syncwarp();
            n_mv += __shfl_sync(FULL_MASK, i_tok, 31);
            ref_base += __shfl_sync(FULL_MASK, i_ref, 31);
            b_base += __shfl_sync(FULL_MASK, i_b, 31);
        }
        if (__any_sync(FULL_MASK, bad)) status = -10;  // "invalid MD", blockjoin.c:1622 / assert :1616
    }
    (void)sm;
    __syncwarp();

    // ---------------- C. vote walk ----------------
    // One known variant per lane.  What the reference's walk over the merged, sorted buffer does at a known entry
    // depends only on its neighbours in that order, and those are found by binary search in the two position-sorted
    // read-variant lists; known variants that share a position are consumed in pairs by the walk (`j += 2`), so of a
    // run of equal positions only an unpaired last one is evaluated.
    int cnt0 = 0, cnt1 = 0;
    uint8_t tag = 254;
    if (status == 0 && P.n_known > 0) {
        const uint32_t kf = P.known_first[ri];
        // the known variants in front of the read's end, and whether they are in position order
        uint32_t ke = kf;
        bool sorted = true;
        for (;;) {
            const uint32_t i = ke + lane;
            const uint32_t pv = i < P.n_known ? P.known[i].pos : 0xffffffffu;
            const bool stop = i >= P.n_known || pv >= end_pos;
            const unsigned sm_ = __ballot_sync(FULL_MASK, stop);
            const uint32_t n_in = sm_ ? (uint32_t)__ffs((int)sm_) - 1u : 32u;
            uint32_t prev = __shfl_up_sync(FULL_MASK, pv, 1);
            if (lane == 0) prev = ke > kf ? P.known[ke - 1].pos : 0u;
            if (__any_sync(FULL_MASK, lane < n_in && pv < prev)) sorted = false;
            ke += n_in;
            if (sm_) break;
        }
        const uint32_t m = ke - kf;
        if (sorted) {
            for (uint32_t jb = 0; jb < m; jb += 32) {
                const uint32_t j = jb + lane;
                if (j >= m) continue;
                const KnownVar kv = P.known[kf + j];
                const uint32_t p = kv.pos;
                // position of this entry inside its run of equal positions
                uint32_t t = 0;
                while (t < j && P.known[kf + j - t - 1].pos == p) t++;
                const bool has_next_known = j + 1 < m;
                const uint32_t nk_pos = has_next_known ? P.known[kf + j + 1].pos : 0xffffffffu;
                if ((t & 1u) || (has_next_known && nk_pos == p)) continue;  // consumed as one of a pair
                const uint32_t ia = lower_bound_u32(ins_ref, n_ins, p), ib = lower_bound_u32(mv_pos, n_mv, p);
                const bool has_a = ia < n_ins, has_b = ib < n_mv;
                if (!has_next_known && !has_a && !has_b) { if (kv.haptag & 1) cnt1++; else cnt0++; continue; }  // last entry of the merged list
                uint32_t rp = 0xffffffffu;
                bool from_ins = false;
                if (has_a) { rp = ins_ref[ia]; from_ins = true; }
                if (has_b && mv_pos[ib] < rp) { rp = mv_pos[ib]; from_ins = false; }
                const bool next_is_read = (has_a || has_b) && !(has_next_known && nk_pos <= rp);
                if (next_is_read && rp == p) {
                    // does the read carry the ALT allele?  (length and bases; the op type is not compared)
                    bool ok;
                    if (from_ins) {
                        ok = kv.len == ins_len[ia];
                        for (uint32_t u = 0; ok && u < kv.len; u++) {
                            const uint32_t sp = ins_self[ia] + u;
                            const uint8_t bch = sp < R.l_qseq ? nt4_of_nib(seq_nib(seq, sp)) : 4;
                            if (P.bases[kv.bases_off + u] != bch) ok = false;
                        }
                    } else {
                        const uint32_t info = mv_info[ib], vlen = info >> 1;
                        ok = kv.len == vlen;
                        if (ok) {
                            if (info & 1u) {
                                for (uint32_t u = 0; ok && u < vlen; u++)
                                    if (P.bases[kv.bases_off + u] != nt4_of_char(md[mv_aux[ib] + u])) ok = false;
                            } else ok = P.bases[kv.bases_off] == (uint8_t)mv_aux[ib];
                        }
                    }
                    if (ok) { if ((kv.haptag ^ 1) & 1) cnt1++; else cnt0++; }
                    continue;
                }
                // next entry sits on another position: REF, unless the previous entry is a read deletion
                // reaching this position (blockjoin.c:1765-1784)
                bool skip = false;
                {
                    const bool has_pa = ia > 0, has_pb = ib > 0;
                    uint32_t ppos = 0;
                    bool prev_is_md = false, have_prev_read = false;
                    if (has_pa) { ppos = ins_ref[ia - 1]; have_prev_read = true; }
                    if (has_pb && (!has_pa || mv_pos[ib - 1] >= ppos)) { ppos = mv_pos[ib - 1]; prev_is_md = true; have_prev_read = true; }
                    // the element before the known entry is a read variant unless a known entry sorts later
                    bool prev_known_later = false;
                    if (j > 0) {
                        const uint32_t pkp = P.known[kf + j - 1].pos;
                        if (!have_prev_read || pkp > ppos) prev_known_later = true;
                    }
                    if (have_prev_read && !prev_known_later && prev_is_md) {
                        const uint32_t info = mv_info[ib - 1];
                        if ((info & 1u) && ppos + (info >> 1) >= p) skip = true;
                    }
                }
                if (!skip) { if (kv.haptag & 1) cnt1++; else cnt0++; }
            }
            cnt0 = warp_sum(cnt0);
            cnt1 = warp_sum(cnt1);
        } else if (lane == 0) {
            // known variants out of position order inside the read's span (rare): the sequential walk over an
            // explicitly sorted index list
            if (m > HAP_MAX_UNSORTED) status = -8;
            else {
                uint32_t order[HAP_MAX_UNSORTED];
                for (uint32_t a = 0; a < m; a++) {  // insertion sort by (pos, index)
                    uint32_t x = kf + a, b = a;
                    while (b > 0 && P.known[order[b - 1]].pos > P.known[x].pos) { order[b] = order[b - 1]; b--; }
                    order[b] = x;
                }
                for (uint32_t j = 0; j < m;) {
                    const uint32_t vi = order[j];
                    const KnownVar kv = P.known[vi];
                    const uint32_t p = kv.pos;
                    const uint32_t ia = lower_bound_u32(ins_ref, n_ins, p), ib = lower_bound_u32(mv_pos, n_mv, p);
                    const bool has_a = ia < n_ins, has_b = ib < n_mv;
                    const bool has_next_known = j + 1 < m;
                    if (!has_next_known && !has_a && !has_b) { if (kv.haptag & 1) cnt1++; else cnt0++; break; }
                    const uint32_t nk_pos = has_next_known ? P.known[order[j + 1]].pos : 0xffffffffu;
                    if (has_next_known && nk_pos == p) { j += 2; continue; }
                    uint32_t rp = 0xffffffffu;
                    bool from_ins = false;
                    if (has_a) { rp = ins_ref[ia]; from_ins = true; }
                    if (has_b && mv_pos[ib] < rp) { rp = mv_pos[ib]; from_ins = false; }
                    const bool next_is_read = (has_a || has_b) && !(has_next_known && nk_pos <= rp);
                    if (next_is_read && rp == p) {
                        bool ok;
                        if (from_ins) {
                            ok = kv.len == ins_len[ia];
                            for (uint32_t u = 0; ok && u < kv.len; u++) {
                                const uint32_t sp = ins_self[ia] + u;
                                const uint8_t bch = sp < R.l_qseq ? nt4_of_nib(seq_nib(seq, sp)) : 4;
                                if (P.bases[kv.bases_off + u] != bch) ok = false;
                            }
                        } else {
                            const uint32_t info = mv_info[ib], vlen = info >> 1;
                            ok = kv.len == vlen;
                            if (ok) {
                                if (info & 1u) {
                                    for (uint32_t u = 0; ok && u < vlen; u++)
                                        if (P.bases[kv.bases_off + u] != nt4_of_char(md[mv_aux[ib] + u])) ok = false;
                                } else ok = P.bases[kv.bases_off] == (uint8_t)mv_aux[ib];
                            }
                        }
                        if (ok) { if ((kv.haptag ^ 1) & 1) cnt1++; else cnt0++; }
                        j += 1;
                        continue;
                    }
                    bool skip = false;
                    {
                        const bool has_pa = ia > 0, has_pb = ib > 0;
                        uint32_t ppos = 0;
                        bool prev_is_md = false, have_prev_read = false;
                        if (has_pa) { ppos = ins_ref[ia - 1]; have_prev_read = true; }
                        if (has_pb && (!has_pa || mv_pos[ib - 1] >= ppos)) { ppos = mv_pos[ib - 1]; prev_is_md = true; have_prev_read = true; }
                        bool prev_known_later = false;
                        if (j > 0) {
                            const uint32_t pkp = P.known[order[j - 1]].pos;
                            if (!have_prev_read || pkp > ppos) prev_known_later = true;
                        }
                        if (have_prev_read && !prev_known_later && prev_is_md) {
                            const uint32_t info = mv_info[ib - 1];
                            if ((info & 1u) && ppos + (info >> 1) >= p) skip = true;
                        }
                    }
                    if (!skip) { if (kv.haptag & 1) cnt1++; else cnt0++; }
                    j += 1;
                }
            }
        }
        status = __shfl_sync(FULL_MASK, status, 0);
        cnt0 = __shfl_sync(FULL_MASK, cnt0, 0);
        cnt1 = __shfl_sync(FULL_MASK, cnt1, 0);
        if (status == 0) {
            const float mx = (float)(cnt0 > cnt1 ? cnt0 : cnt1);
            const int mn = cnt0 <= cnt1 ? cnt0 : cnt1;
            const float ratio = mn == 0 ? 0.f : __fdiv_rn(mx, (float)mn);
            if ((cnt0 > 3 && cnt1 > 3 && ratio < 5.f) || cnt0 == cnt1) tag = 254;
            else tag = cnt0 > cnt1 ? 0 : 1;
        }
    }
    if (lane == 0) {
        P.out_tag[ri] = tag;
        P.out_status[ri] = status;
        P.out_votes[2 * ri] = cnt0;
        P.out_votes[2 * ri + 1] = cnt1;
        if (P.out_counts) { P.out_counts[2 * ri] = n_ins; P.out_counts[2 * ri + 1] = status == 0 ? n_mv : 0u; }
    }
}

// ---------------------------------------------------------------------------------------------
// Votes for the phase of unphased variants inside a dropped interval (recover_variant_phase_in_one_interval,
// reference blockjoin.c:2475-2600): every read that methylation phasing tagged contributes, at each known variant
// position where the read itself shows a variant (an insertion from its CIGAR, a mismatch or deletion from its MD —
// the lists parse_variants_for_one_read builds, left in the scratch arrays by haptag_kernel), one vote for its
// haplotype.  One warp per read; positions are looked up in the sorted list of known positions.
// votes: [2 * n_pos] counts per position and haplotype, then [2 * n_pos] = read variants at or behind the last
// known position (the reference never evaluates a known variant that ends its merged list, :2561).
// ---------------------------------------------------------------------------------------------
struct VariantVoteParams {
    uint32_t n_reads;
    const uint32_t *ins_off, *mdv_off;
    const uint32_t *ins_ref, *mv_pos;
    const uint32_t *counts;     // 2 per read (haptag_kernel out_counts)
    const uint8_t *read_hap;    // per read: haplotype from methylation phasing, 255 = the read does not take part
    const uint32_t *poss;       // known positions, ascending
    uint32_t n_pos;
    int32_t *votes;
};

__global__ void __launch_bounds__(HAP_WARPS * 32) variant_vote_kernel(VariantVoteParams P) {
    const uint32_t ri = blockIdx.x * HAP_WARPS + (threadIdx.x >> 5);
    if (ri >= P.n_reads || P.n_pos == 0) return;
    const uint32_t hap = P.read_hap[ri];
    if (hap == 255u) return;
    const unsigned lane = lane_id();
    const uint32_t last = P.poss[P.n_pos - 1];
    for (int list = 0; list < 2; list++) {
        const uint32_t *v = list == 0 ? P.ins_ref + P.ins_off[ri] : P.mv_pos + P.mdv_off[ri];
        const uint32_t n = P.counts[2 * ri + list];
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t p = v[i];
            if (p >= last) atomicAdd(&P.votes[2 * P.n_pos], 1);
            if (hap > 1u) continue;
            // every known variant at this position (duplicates are adjacent)
            for (uint32_t k = lower_bound_u32(P.poss, P.n_pos, p); k < P.n_pos && P.poss[k] == p; k++) atomicAdd(&P.votes[2 * k + hap], 1);
        }
    }
}

}  // namespace pomfret_gpu
#endif
