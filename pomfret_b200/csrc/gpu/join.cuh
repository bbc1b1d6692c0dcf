// (e) Greedy haplotag propagation across a phase-block gap, one CTA per (window, direction).
//
// Replaces haplotag_region1 (reference blockjoin.c:3958-4080) with everything it calls:
//   insert_ref_reads_methmer_counts 3776-3810, insert_mmrs_to_counts 3453-3486,
//   query_counts_of_mmrs 3487-3515, use_mmr_count_predict_tag_for_one_read 3594-3656,
//   update_available_methmer_range 3669-3691, predict_tags_of_reads 3693-3774,
// and tabulates the 2x2 table of evaluate_separation1 (3881-3893).  Fisher's test and the join rule
// stay on the host (fp64 lgamma, a handful of integers).
//
// The reference keeps, per site, a growing list of (key, count per haplotype); membership is all that
// the list adds over a dense table, and "key present" == "some haplotype count is non-zero", so the
// table here is dense: one 32-bit word per (site, key) holding both 16-bit counts, one word per site
// for the two sums.  It lives in the CTA's shared memory (about 100 KB for a 30x window at k = 3; windows
// whose tables exceed an SM fall back to a global pool), next to the per-read methmer metadata, the scan
// order, a tagged bitmap and the candidates' methmer keys.
//
// Score sums are order sensitive fp32 (blockjoin.c:3620-3636): values are produced in parallel, one
// lane per methmer (IEEE divides), the non-zero ones compacted in methmer order and then added strictly
// in that order with round-to-nearest adds (no fast-math, no FMA contraction possible).
//
// The loop is latency bound (one read tagged per iteration), so everything on its critical path is kept
// short: candidates live in fixed slots, the next candidate and its keys are prefetched by a service warp
// while the others score, two barriers per iteration.
#ifndef POMFRET_GPU_JOIN_CUH
#define POMFRET_GPU_JOIN_CUH
#include "gpu_rt.h"
#include "types.h"

namespace pomfret_gpu {

#ifndef POMFRET_JOIN_THREADS
#define POMFRET_JOIN_THREADS 512
#endif
constexpr int JOIN_THREADS = POMFRET_JOIN_THREADS;
constexpr int JOIN_WARPS = JOIN_THREADS / 32;
constexpr int JOIN_MAX_CAND = 1024;       // candidate slots are sized at launch (dynamic shared memory); this only bounds the request
constexpr int JOIN_U = 8;                 // methmer sub-chunks (32 each) whose loads are issued together
constexpr int JOIN_CHUNK = JOIN_U * 32;   // methmers of a candidate whose keys are cached in shared memory
constexpr int JOIN_STAGE = 128;           // score terms staged per warp between two runs of the ordered sum

struct JoinParams {
    const WindowRec *win;
    WindowState *state;
    const uint32_t *rs_rev;
    const int32_t *rs_hp;  // initial tags per slot
    const uint32_t *ids_left, *ids_left_strict, *ids_right, *ids_right_strict;
    const uint32_t *site_pos;
    const uint32_t *mm_off[2], *mm_n[2], *mm_start[2];
    const uint32_t *mmr_pool;
    uint32_t *tab;         // table pool (used by windows whose tables do not fit shared memory)
    uint8_t *tags[2];      // propagated tags per slot and direction
    uint32_t *order[2];    // tagging order per direction (slot-indexed slices)
    int32_t n_cand;
    int32_t cov_run;
    int32_t k;
    const uint32_t *cta_map;  // blockIdx.x -> window * 2 + direction
    uint32_t smem_tab_words;  // shared-memory words reserved for the count tables
    uint32_t meta_cap;        // reads whose per-read state fits the shared-memory arrays
    uint32_t stage_cap;       // score terms a candidate slot can stage before its warp has to fold them into the running sums
};

// Count tables: per site one row of 3^k key words (lo16 = haplotype 0 count, hi16 = haplotype 1 count)
// plus one word with the two sums.  Methmer symbols are m=0, u=1, -=2 (blockjoin.c:3186-3194), so a key's
// base-4 digits never contain 3 and the row is indexed by the same digits read in base 3.  The row stride
// is odd so that consecutive sites fall into different shared-memory banks.
__host__ __device__ __forceinline__ uint32_t join_n_keys(int k) { uint32_t v = 1; for (int i = 0; i < k; i++) v *= 3u; return v; }
__host__ __device__ __forceinline__ uint32_t join_row_stride(int k) { return (join_n_keys(k) + 1u) | 1u; }
__device__ __forceinline__ uint32_t compact_key(uint32_t key) {  // k <= 8: the base-4 digits of the methmer key read in base 3
    uint32_t v = ((key >> 6) & 3u) * 27u + ((key >> 4) & 3u) * 9u + ((key >> 2) & 3u) * 3u + (key & 3u);
    if (key >> 8) v += (((key >> 14) & 3u) * 27u + ((key >> 12) & 3u) * 9u + ((key >> 10) & 3u) * 3u + ((key >> 8) & 3u)) * 81u;
    return v;
}
// dynamic shared memory of join_kernel for the given capacities (bytes)
// Score terms staged per candidate slot.  Measured on the 60x batch (profiles/r02_join_variants.txt): the kernel is
// issue bound once two or three CTAs share an SM, so shared memory is better spent on a third CTA than on rows that
// hold every term of a read: 12 KB for all slots (a row that runs full is folded early by its own warp).
#ifndef POMFRET_JOIN_STAGE_KB
#define POMFRET_JOIN_STAGE_KB 12
#endif
#ifndef POMFRET_JOIN_MIN_CTAS
#define POMFRET_JOIN_MIN_CTAS 3
#endif
__host__ __device__ __forceinline__ uint32_t join_stage_cap(int n_cand) {
    const uint32_t fit = (uint32_t)(POMFRET_JOIN_STAGE_KB * 1024) / ((uint32_t)(n_cand + 1) * 8u);
    const uint32_t cap = fit > (uint32_t)JOIN_CHUNK + 1u ? (uint32_t)JOIN_CHUNK : (fit > 34u ? fit - 2u : 32u);
    return cap & ~1u;  // even, so that the row stride cap + 1 (in 8-byte terms) is odd: lanes of the summing warp hit distinct banks
}
__host__ __device__ __forceinline__ size_t join_smem_bytes(uint32_t tab_entries, uint32_t entry_bytes, uint32_t meta_cap, int n_cand, int n_warps) {
    (void)n_warps;
    return (((size_t)tab_entries * entry_bytes + 7) & ~(size_t)7) + (size_t)(n_cand + 1) * (join_stage_cap(n_cand) + 1) * 8 +
           (size_t)meta_cap * 12 + (size_t)((meta_cap + 1) / 2) * 4 + (size_t)((meta_cap + 31) / 32) * 4 +
           (size_t)(n_cand + 1) * JOIN_CHUNK + (size_t)(n_cand + 2) * 48 + 64;  // 48: the per-slot state arrays
}

// A table entry holds the counts of both haplotypes: 16 + 16 bits in a 32-bit word, or 8 + 8 bits in a 16-bit word for
// windows in which fewer than 256 reads touch any one site (WindowState::max_cov): half the shared memory per CTA,
// twice the CTAs per SM.  Insertions by the extension loop go to distinct entries (one read, one methmer per site) and are
// plain read-modify-writes; the seeding phase adds from several warps at once and uses atomics on the enclosing word.
template <typename TabT> struct JoinTab;
template <> struct JoinTab<uint32_t> {
    static constexpr uint32_t kShift = 16, kMask = 0xffffu;
    static __device__ __forceinline__ void atomic_add(uint32_t *p, uint32_t inc) { atomicAdd(p, inc); }
};
template <> struct JoinTab<uint16_t> {
    static constexpr uint32_t kShift = 8, kMask = 0xffu;
    static __device__ __forceinline__ void atomic_add(uint16_t *p, uint32_t inc) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        atomicAdd(reinterpret_cast<uint32_t *>(a & ~(uintptr_t)3), inc << ((a & 2u) * 8u));  // (no count reaches 256: no carry)
    }
};

// Range growth of update_available_methmer_range (blockjoin.c:3669-3691), warp-parallel: the left edge
// walks down from `mn` while the site's coverage reaches cov, the right edge walks up from `mx`.
template <typename TabT>
__device__ __forceinline__ void grow_range(const TabT *tab, uint32_t stride, uint32_t n_keys, uint32_t n_sites, int cov,
                                           uint32_t &mn, uint32_t &mx) {
    constexpr uint32_t kShift = JoinTab<TabT>::kShift, kMask = JoinTab<TabT>::kMask;
    const int lane = (int)lane_id();
    for (int i0 = (int)mn;; i0 -= 32) {
        const int i = i0 - lane;
        bool ok = false;
        if (i >= 0) {
            const uint32_t sm = tab[(size_t)i * stride + n_keys];
            ok = (int)((sm & kMask) + (sm >> kShift)) >= cov;
        }
        const unsigned bad = ~__ballot_sync(FULL_MASK, ok);  // first lane that stops the walk (or ran past site 0)
        const int f = bad ? __ffs((int)bad) - 1 : 32;
        if (f > 0) mn = (uint32_t)(i0 - (f - 1));
        if (f < 32) break;
    }
    for (int i0 = (int)mx;; i0 += 32) {
        const int i = i0 + lane;
        bool ok = false;
        if (i < (int)n_sites) {
            const uint32_t sm = tab[(size_t)i * stride + n_keys];
            ok = (int)((sm & kMask) + (sm >> kShift)) >= cov;
        }
        const unsigned bad = ~__ballot_sync(FULL_MASK, ok);
        const int f = bad ? __ffs((int)bad) - 1 : 32;
        if (f > 0) mx = (uint32_t)(i0 + (f - 1));
        if (f < 32) break;
    }
}

// kTabSmem: every window of the launch keeps its count tables in shared memory (the compiler then emits
// shared-memory loads/stores for them); otherwise windows that do not fit use the global pool.
template <bool kTabSmem, typename TabT> __global__ void __launch_bounds__(JOIN_THREADS, POMFRET_JOIN_MIN_CTAS) join_kernel(JoinParams P) {
    constexpr uint32_t kShift = JoinTab<TabT>::kShift, kMask = JoinTab<TabT>::kMask;
    __shared__ int s_i_last, s_failed, s_done, s_fill, s_cursor, s_next_id;
    __shared__ int s_newest[2], s_nocc[2];
    __shared__ uint32_t s_min, s_max, s_seq;
    // candidate slots, two copies: an iteration reads one and writes the other (no barrier between the warps that
    // still look for the best candidate and the one that already recycles its slot)
    __shared__ int s_best;
    __shared__ int s_tbl[4];
    POMFRET_DYN_SMEM(uint32_t, dyn);

    const uint32_t wd = P.cta_map[blockIdx.x];
    const uint32_t w = wd >> 1, d = wd & 1u;
    const WindowRec W = P.win[w];
    WindowState &S = P.state[w];
    const uint32_t n = S.n, n_sites = S.n_sites;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const uint32_t nthreads = blockDim.x, nwarps = nthreads >> 5;
    if (n == 0 || n_sites == 0 || S.status != 0) return;
    const uint32_t first = W.first_read;
    const uint32_t n_keys = join_n_keys(P.k);
    const uint32_t stride = join_row_stride(P.k);
    uint8_t *tags = P.tags[d] + first;
    const uint32_t *g_off = P.mm_off[d] + first, *g_n = P.mm_n[d] + first, *g_start = P.mm_start[d] + first;
    const uint32_t *pool = P.mmr_pool;
    const uint32_t *site_pos = P.site_pos + W.site_off;
    const uint32_t *ref_ids = (d == 0 ? P.ids_left : P.ids_right) + first;
    const uint32_t *rev = P.rs_rev + first;
    const uint32_t n_ref = d == 0 ? S.n_left : S.n_right;
    const int n_cand = P.n_cand;
    const bool key_fits_u8 = P.k <= 5;  // 3^5 = 243 compact keys: the candidates' key cache holds one byte per methmer

    // ---- shared-memory carve-up ----
    TabT *s_tab = reinterpret_cast<TabT *>(dyn);
    float2 *s_stage = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(dyn) + (((size_t)P.smem_tab_words * sizeof(TabT) + 7) & ~(size_t)7));  // [nwarps][JOIN_STAGE]
    const uint32_t stage_cap = P.stage_cap, stage_stride = P.stage_cap + 1u;  // [n_cand + 1][stage_stride]
    uint32_t *s_moff = reinterpret_cast<uint32_t *>(s_stage + (size_t)(P.n_cand + 1) * stage_stride);  // per read: offset of its keys in the pool
    uint32_t *s_mn = s_moff + P.meta_cap;                      // per read: number of methmers
    uint32_t *s_mst = s_mn + P.meta_cap;                       // per read: site index of the first one
    uint16_t *s_scan = reinterpret_cast<uint16_t *>(s_mst + P.meta_cap);  // scan order -> read id (direction 1)
    uint32_t *s_tagged = reinterpret_cast<uint32_t *>(s_scan) + (P.meta_cap + 1) / 2;  // bit per read: tagged 0/1
    uint8_t *s_keys = reinterpret_cast<uint8_t *>(s_tagged + (P.meta_cap + 31) / 32);   // [n_cand + 1][JOIN_CHUNK]
    // per-slot state, n_cand + 1 entries each (the first array holds 8-byte items: its address is rounded up)
    const uint32_t n_sl = (uint32_t)P.n_cand + 1u, n_sl2 = (n_sl + 1u) & ~1u;
    float2 *s_pre = reinterpret_cast<float2 *>((reinterpret_cast<uintptr_t>(s_keys + (size_t)n_sl * JOIN_CHUNK) + 7) & ~(uintptr_t)7);  // ordered sums of the terms folded before (rows that ran full)
    int2 *s_ll = reinterpret_cast<int2 *>(s_pre + n_sl);                      // the two score_h_l counters of the slot
    uint32_t *s_sid_[2], *s_sseq_[2];
    s_sid_[0] = reinterpret_cast<uint32_t *>(s_ll + n_sl);                   // read id
    s_sid_[1] = s_sid_[0] + n_sl2;
    s_sseq_[0] = s_sid_[1] + n_sl2;                                          // position in scan order (monotone counter), SLOT_FREE if empty
    s_sseq_[1] = s_sseq_[0] + n_sl2;
    int *s_tag = reinterpret_cast<int *>(s_sseq_[1] + n_sl2);
    uint32_t *s_nz = reinterpret_cast<uint32_t *>(s_tag + n_sl2);            // score terms staged in the slot's row
#define s_sid(b) s_sid_[b]
#define s_sseq(b) s_sseq_[b]
    const bool tab_in_smem = kTabSmem || (size_t)n_sites * stride <= P.smem_tab_words;
    const bool meta_in_smem = n <= P.meta_cap;
    TabT *tab = kTabSmem ? s_tab : (tab_in_smem ? s_tab : reinterpret_cast<TabT *>(P.tab) + (size_t)S.tab_base[d] * stride);

    // ---- wipe the tables (insert_ref_reads_methmer_counts, :3780-3789), load the per-read state ----
    for (size_t i = tid, words = (size_t)n_sites * stride; i < words; i += nthreads) tab[i] = 0;
    if (meta_in_smem) {
        for (uint32_t i = tid; i < n; i += nthreads) {
            s_moff[i] = g_off[i]; s_mn[i] = g_n[i]; s_mst[i] = g_start[i];
            s_scan[i] = (uint16_t)(d == 0 ? i : rev[i]);
        }
        for (uint32_t i = tid; i < (n + 31) / 32; i += nthreads) s_tagged[i] = 0;
    }
    // ---- available range, :3976-4004 (warp-parallel count of the sites on the starting side of the gap) ----
    if (warp == 0) {
        uint32_t mn, mx;
        if (d == 0) {
            mn = 0; mx = 0;
            for (uint32_t i0 = 0; i0 < n_sites; i0 += 32) {
                const uint32_t i = i0 + lane;
                const unsigned ok = __ballot_sync(FULL_MASK, i < n_sites && site_pos[i] <= W.ref_start);
                if (ok == FULL_MASK) { mx += 32; continue; }
                mx += (uint32_t)__ffs((int)~ok) - 1u;
                break;
            }
        } else {
            mn = n_sites - 1; mx = n_sites - 1;
            for (int i0 = (int)n_sites - 1; i0 >= 0; i0 -= 32) {
                const int i = i0 - (int)lane;
                const unsigned ok = __ballot_sync(FULL_MASK, i >= 0 && site_pos[i] > W.ref_end);
                if (ok == FULL_MASK) { mn -= 32; continue; }
                mn -= (uint32_t)__ffs((int)~ok) - 1u;
                break;
            }
        }
        if (lane == 0) {
            s_min = mn; s_max = mx;
            s_failed = 0; s_done = 0; s_fill = -1;
            s_i_last = d == 0 ? 0 : (int)n - 1;
            s_tbl[0] = s_tbl[1] = s_tbl[2] = s_tbl[3] = 0;
        }
    }
    __syncthreads();
    // ---- seed with the reference reads of the starting side, :3793-3803 ----
    // (a read touches each site at most once and the 16-bit halves add independently: order is irrelevant)
    for (uint32_t r = warp; r < n_ref; r += nwarps) {
        const uint32_t id = ref_ids[r];
        const int hap = P.rs_hp[first + id];
        if (hap == 0 || hap == 1) {
            const uint32_t nm = g_n[id], st = g_start[id], off = g_off[id];
            const uint32_t inc = hap == 0 ? 1u : 1u << kShift;
            for (uint32_t i0 = lane; i0 < nm; i0 += 32) {
                TabT *row = tab + (size_t)(st + i0) * stride;
                JoinTab<TabT>::atomic_add(&row[compact_key(pool[off + i0])], inc);
                JoinTab<TabT>::atomic_add(&row[n_keys], inc);
            }
        }
    }
    // ---- un-tag everything but the reference reads, :4010-4025 (with the (id<<2)|hp packing) ----
    for (uint32_t i = tid; i < n; i += nthreads) tags[i] = 2;
    __syncthreads();
    // The reference re-tags the list sequentially, and with hp = 254 the packing corrupts the id (the write
    // lands on read id | 63): a later entry may overwrite an earlier one.  One warp walks the list in order,
    // 32 entries at a time; inside a group the highest lane writing to a read wins.
    if (warp == 0) {
        for (uint32_t r0 = 0; r0 < n_ref; r0 += 32) {
            const uint32_t r = r0 + lane;
            uint32_t tid2 = 0xffffffffu - lane;  // distinct dummies for idle lanes
            uint8_t t = 0;
            bool act = false;
            if (r < n_ref) {
                const uint32_t id = ref_ids[r];
                const uint32_t packed = (id << 2) | (uint32_t)P.rs_hp[first + id];
                if ((packed >> 2) < n) { act = true; tid2 = packed >> 2; t = (uint8_t)(packed & 3u); }
            }
            const unsigned same = __match_any_sync(FULL_MASK, tid2);
            if (act && (same >> lane) <= 1u) {  // no higher lane targets the same read
                tags[tid2] = t;
                if (meta_in_smem) {
                    if (t < 2) atomicOr(&s_tagged[tid2 >> 5], 1u << (tid2 & 31u));
                    else atomicAnd(&s_tagged[tid2 >> 5], ~(1u << (tid2 & 31u)));
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();

    auto is_untagged = [&](uint32_t id) -> bool {
        if (meta_in_smem) return !((s_tagged[id >> 5] >> (id & 31u)) & 1u);
        const uint8_t t = tags[id];
        return t != 0 && t != 1;
    };
    // Candidates ("the first n_cand untagged reads in scan order from i_last", :4039-4045) live in n_cand + 1
    // fixed slots, each with its read id, its position in scan order (a running counter: ties between equal
    // scores go to the later candidate, :3729-3760) and a cache of its methmer keys.  The extra slot holds the
    // look-ahead entry: the next untagged read behind the scan cursor, found and fetched from global memory by
    // the last warp while the others score, one iteration before it is first scored.  A tagged read simply
    // frees its slot for the next look-ahead entry; only a failure (i_last moves, :4064-4068) rebuilds the slots.
    constexpr uint32_t SLOT_FREE = 0xffffffffu;
    const int n_slots = n_cand + 1;
    auto scan_id = [&](int i0) -> uint32_t { return d == 0 ? (uint32_t)i0 : (meta_in_smem ? (uint32_t)s_scan[i0] : rev[i0]); };
    auto fill_keys = [&](int buf, int slot) {  // one warp: keys of the slot's read, global -> shared (compact u8)
        const uint32_t id = s_sid(buf)[slot];
        const uint32_t nm = meta_in_smem ? s_mn[id] : g_n[id], off = meta_in_smem ? s_moff[id] : g_off[id];
        if (nm <= JOIN_CHUNK && key_fits_u8) {  // longer reads (and keys of more than five symbols) are scored straight from the pool
            uint8_t *dk = s_keys + (size_t)slot * JOIN_CHUNK;
            for (uint32_t i = lane; i < nm; i += 32) dk[i] = (uint8_t)compact_key(pool[off + i]);
        }
    };
    // next untagged read at or behind scan position `cursor` (one warp); returns its id or -1, moves the cursor behind it
    auto next_untagged = [&](int &cursor) -> int {
        for (;;) {
            const int i0 = d == 0 ? cursor + (int)lane : cursor - (int)lane;
            const bool in = d == 0 ? i0 < (int)n : i0 >= 0;
            uint32_t id = 0;
            bool unt = false;
            if (in) { id = scan_id(i0); unt = is_untagged(id); }
            const unsigned um = __ballot_sync(FULL_MASK, unt);
            if (um) {
                const int fl = __ffs((int)um) - 1;
                cursor += d == 0 ? fl + 1 : -(fl + 1);
                return (int)__shfl_sync(FULL_MASK, id, fl);
            }
            cursor += d == 0 ? 32 : -32;
            if (__ballot_sync(FULL_MASK, in) != FULL_MASK) return -1;  // ran past the last read
        }
    };
    // (re)build all slots from i_last (warp 0): keys are fetched at once, the look-ahead entry included
    auto rebuild_slots = [&](int buf) {
        int cursor = s_i_last, nocc = 0;
        uint32_t seq = 0;
        for (int sl = lane; sl < n_slots; sl += 32) s_sseq(buf)[sl] = SLOT_FREE;
        __syncwarp();
        while (nocc < n_slots) {
            const int id = next_untagged(cursor);
            if (id < 0) break;
            if (lane == 0) { s_sid(buf)[nocc] = (uint32_t)id; s_sseq(buf)[nocc] = seq; }
            __syncwarp();
            fill_keys(buf, nocc);
            nocc++; seq++;
        }
        if (lane == 0) {
            s_nocc[buf] = nocc; s_seq = seq; s_cursor = cursor; s_fill = -1;
            s_newest[buf] = nocc == n_slots ? n_slots - 1 : -1;
            s_next_id = -2;  // unknown: the last warp looks for it while the others score
        }
        __syncwarp();
    };

    // ---- extension loop, :4032-4071 ----
    // Two barriers per iteration.  Between them: (1) every warp grows the available range on its own (the
    // walk is idempotent) and scores the candidate of its slot; the last warp also serves the look-ahead slot;
    // (2) every warp finds the best candidate on its own, all threads insert its methmers, one thread updates
    // the slot.
#ifdef POMFRET_JOIN_PROF
    long long pf_t0 = clock64(), pf[6] = {0, 0, 0, 0, 0, 0}, pf_t = pf_t0;
    int pf_iter = 0;
#define PF_MARK(i) do { long long now_ = clock64(); pf[i] += now_ - pf_t; pf_t = now_; } while (0)
#else
#define PF_MARK(i) do {} while (0)
#endif
    uint32_t n_order = 0;
    int cur = 0;
    bool grow = true;  // update_available_methmer_range after seeding / after every insertion (:3770, :3806)
    if (warp == 0) {
        const int i_last = s_i_last;
        if ((d == 0 && i_last >= (int)n) || (d != 0 && i_last <= 0)) { if (lane == 0) s_done = 1; }
        else rebuild_slots(0);
    }
    __syncthreads();
    for (;;) {
        if (s_done) break;
        const int nocc = s_nocc[cur], newest = s_newest[cur];
        const bool full = nocc == n_slots;
        uint32_t rmin = s_min, rmax = s_max;
        if (grow) {
            grow_range<TabT>(tab, stride, n_keys, n_sites, P.cov_run, rmin, rmax);
            if (tid == 0) { s_min = rmin; s_max = rmax; }  // others may still read the old pair: growing again is harmless
        }
        PF_MARK(0);  // grow
        // ---- score terms of the candidates, one warp per slot (use_mmr_count_predict_tag_for_one_read, :3594-3656):
        //      lookups and divisions in parallel, the non-zero terms compacted in methmer order into the slot's row ----
        for (int c = (int)warp; c < n_slots; c += (int)nwarps) {
            if (s_sseq(cur)[c] == SLOT_FREE || (full && c == newest)) { if (lane == 0) s_tag[c] = -2; continue; }  // empty / look-ahead
            const uint32_t id = s_sid(cur)[c];
            const uint32_t nm = meta_in_smem ? s_mn[id] : g_n[id], st = meta_in_smem ? s_mst[id] : g_start[id];
            const uint32_t off = meta_in_smem ? s_moff[id] : g_off[id];
            const bool cached = nm <= JOIN_CHUNK && key_fits_u8;
            const uint8_t *ck = s_keys + (size_t)c * JOIN_CHUNK;
            float sc0 = 0.f, sc1 = 0.f;  // ordered sums of terms folded early (only when the row runs full)
            int l0 = 0, l1 = 0;
            // only methmers whose site lies in the available range [rmin, rmax) are looked up (:3499-3502)
            const uint32_t i_lo = rmin > st ? rmin - st : 0u;
            const uint32_t i_hi = rmax > st ? (rmax - st < nm ? rmax - st : nm) : 0u;
            float2 *stage = s_stage + (size_t)c * stage_stride;
            uint32_t nz = 0;  // non-zero terms staged (adding +0.0f is exact, so zero terms are dropped)
            for (uint32_t base = i_lo; base < i_hi; base += JOIN_STAGE) {
                // up to four sub-chunks of 32 methmers at a time: their lookups and divisions overlap
                uint32_t key[4], cnt[4], sums[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i0 = base + u * 32 + lane;
                    key[u] = i0 < i_hi ? (cached ? (uint32_t)ck[i0] : compact_key(pool[off + i0])) : 0xffffffffu;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    cnt[u] = 0; sums[u] = 0;
                    if (key[u] != 0xffffffffu) {
                        const TabT *row = tab + (size_t)(st + base + u * 32 + lane) * stride;
                        cnt[u] = row[key[u]];
                        sums[u] = row[n_keys];
                    }
                }
                float v0[4], v1[4];
                bool p0[4], p1[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    v0[u] = v1[u] = 0.f; p0[u] = p1[u] = false;
                    if (base + u * 32 < i_hi) {  // uniform
                        const uint32_t sum0 = sums[u] & kMask, sum1 = sums[u] >> kShift;
                        const uint32_t c0 = cnt[u] & kMask, c1 = cnt[u] >> kShift;
                        p0[u] = cnt[u] != 0 && sum0 != 0;  // key present at this site and the haplotype has counts
                        p1[u] = cnt[u] != 0 && sum1 != 0;
                        // (zero operands are replaced by 1 so that the division always takes its fast path)
                        const float q0 = __fdiv_rn((float)(c0 ? c0 : 1u), (float)(sum0 ? sum0 : 1u));
                        const float q1 = __fdiv_rn((float)(c1 ? c1 : 1u), (float)(sum1 ? sum1 : 1u));
                        v0[u] = p0[u] && c0 ? q0 : 0.f;
                        v1[u] = p1[u] && c1 ? q1 : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (base + u * 32 >= i_hi) break;
                    // (the counts behind score_h_l, :3620-3636, are kept per lane and reduced once per candidate)
                    l0 += (int)p0[u] + (int)(v0[u] > 0.f);
                    l1 += (int)p1[u] + (int)(v1[u] > 0.f);
                    const bool nzv = v0[u] > 0.f || v1[u] > 0.f;
                    const unsigned zz = __ballot_sync(FULL_MASK, nzv);
                    if (nz + (uint32_t)__popc(zz) > stage_cap) {
                        // the row is full (a read with more methmers than a row holds): fold what is staged into the running
                        // sums, strictly in order (every lane walks the row: broadcast reads)
                        __syncwarp();
                        for (uint32_t t = 0; t < nz; t++) { const float2 r = stage[t]; sc0 = __fadd_rn(sc0, r.x); sc1 = __fadd_rn(sc1, r.y); }
                        __syncwarp();
                        nz = 0;
                    }
                    if (nzv) stage[nz + __popc(zz & ((1u << lane) - 1u))] = make_float2(v0[u], v1[u]);
                    nz += __popc(zz);
                }
            }
            l0 = (int)__reduce_add_sync(FULL_MASK, (unsigned)l0);
            l1 = (int)__reduce_add_sync(FULL_MASK, (unsigned)l1);
            if (lane == 0) {
                s_nz[c] = nz;
                s_pre[c] = make_float2(sc0, sc1);
                s_ll[c] = int2{l0, l1};
                s_tag[c] = -3;  // terms ready, sums pending
            }
        }
        PF_MARK(1);  // own terms
        if (warp == nwarps - 1) {
            // look-ahead service, off the critical path: keys of the entry placed last iteration, then the read
            // that will take the next freed slot
            if (s_fill >= 0) fill_keys(cur, s_fill);
            if (s_next_id == -2 || s_fill >= 0) {
                int cursor = s_cursor;
                const int nx = full ? next_untagged(cursor) : -1;  // a list that is not full has run out of reads
                if (lane == 0) { s_next_id = nx; s_cursor = cursor; }
            }
        }
        __syncthreads();
        PF_MARK(2);  // wait for the slowest scorer / the look-ahead service
        // ---- ordered sums, one LANE per candidate (warp 0): the additions of blockjoin.c:3620-3636 are a serial chain per
        //      candidate and haplotype, so a whole warp per candidate would spend an issue slot per addition; one warp walks
        //      all rows at once (row stride odd in 8-byte units: no bank conflicts).  Then the best candidate:
        //      stable ascending sort + scan from the top == max score, ties to the later candidate (:3729-3760); scores are
        //      >= 0, so their bit patterns order like unsigned ints ----
        if (warp == 0) {
            uint32_t bs = 0, bq = 0;
            int bc = -1;
            for (int g0 = 0; g0 < n_slots; g0 += 32) {
                const int c = g0 + (int)lane;
                const bool act = c < n_slots && s_tag[c] == -3;
                const uint32_t nz = act ? s_nz[c] : 0u;
                const uint32_t mx = __reduce_max_sync(FULL_MASK, nz);
                const float2 *row = s_stage + (size_t)(c < n_slots ? c : 0) * stage_stride;
                float sc0 = 0.f, sc1 = 0.f;
                if (act) { const float2 pre = s_pre[c]; sc0 = pre.x; sc1 = pre.y; }
                for (uint32_t t0 = 0; t0 < mx; t0 += 8) {
                    float2 r[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) r[j] = row[t0 + j < stage_cap ? t0 + j : stage_cap - 1u];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (t0 + j < nz) { sc0 = __fadd_rn(sc0, r[j].x); sc1 = __fadd_rn(sc1, r[j].y); }
                }
                if (act) {
                    const int2 ll = s_ll[c];
                    const float diff = sc0 > sc1 ? __fsub_rn(sc0, sc1) : __fsub_rn(sc1, sc0);
                    if (!(diff < 3.0f && (ll.x < 3 || ll.y < 3))) {
                        const uint32_t sb = __float_as_uint(diff), sq = s_sseq(cur)[c];
                        s_tag[c] = sc0 > sc1 ? 0 : 1;
                        if (bc < 0 || sb > bs || (sb == bs && sq > bq)) { bs = sb; bq = sq; bc = c; }
                    } else s_tag[c] = -1;
                }
            }
            const uint32_t top = __reduce_max_sync(FULL_MASK, bc >= 0 ? bs : 0u);
            const bool cand = bc >= 0 && bs == top;
            const uint32_t topq = __reduce_max_sync(FULL_MASK, cand ? bq + 1u : 0u);
            const unsigned who = __ballot_sync(FULL_MASK, cand && bq + 1u == topq);
            int best_w0 = -1;
            if (who) best_w0 = __shfl_sync(FULL_MASK, bc, __ffs((int)who) - 1);
            if (lane == 0) s_best = best_w0;
        }
        __syncthreads();
        const int best = s_best;
        const uint32_t best_id = best >= 0 ? s_sid(cur)[best] : 0u;
        const int hap = best >= 0 ? s_tag[best] : -1;
        PF_MARK(3);  // best
        if (best >= 0) {
            // ---- insert_mmrs_to_counts (:3453-3486), all threads ----
            const uint32_t nm = meta_in_smem ? s_mn[best_id] : g_n[best_id], st = meta_in_smem ? s_mst[best_id] : g_start[best_id];
            const uint32_t off = meta_in_smem ? s_moff[best_id] : g_off[best_id];
            const bool cached = nm <= JOIN_CHUNK && key_fits_u8;
            const uint8_t *ck = s_keys + (size_t)best * JOIN_CHUNK;
            const TabT inc = (TabT)(hap == 0 ? 1u : 1u << kShift);
            for (uint32_t i0 = tid; i0 < nm; i0 += nthreads) {
                TabT *row = tab + (size_t)(st + i0) * stride;
                row[cached ? (uint32_t)ck[i0] : compact_key(pool[off + i0])] += inc;
                row[n_keys] += inc;
            }
            if (warp == nwarps - 1) {
                // next iteration's slots: a copy, with the freed slot taking the look-ahead entry found during
                // scoring (its keys follow next iteration)
                const int nx = s_next_id;
                const uint32_t seq_new = s_seq;
                for (int c = (int)lane; c < n_slots; c += 32) {
                    uint32_t sid = s_sid(cur)[c], sq = s_sseq(cur)[c];
                    if (c == best) { if (nx >= 0) { sid = (uint32_t)nx; sq = seq_new; } else sq = SLOT_FREE; }
                    s_sid(cur ^ 1)[c] = sid;
                    s_sseq(cur ^ 1)[c] = sq;
                }
                __syncwarp();
                if (lane == 31) {
                    tags[best_id] = (uint8_t)hap;
                    if (meta_in_smem) s_tagged[best_id >> 5] |= 1u << (best_id & 31u);
                    P.order[d][first + n_order] = best_id;
                    s_failed = 0;
                    if (nx >= 0) { s_seq = seq_new + 1; s_nocc[cur ^ 1] = nocc; s_newest[cur ^ 1] = best; s_fill = best; }
                    else { s_nocc[cur ^ 1] = nocc - 1; s_newest[cur ^ 1] = -1; s_fill = -1; }
                }
            }
            n_order++;
        } else if (warp == 0) {
            // ---- nothing could be tagged: move i_last (:4064-4068) and start over from there ----
            const int fl = s_failed + 1, il = s_i_last + (d == 0 ? n_cand : -n_cand);
            const bool done = fl > 10 || (d == 0 && il >= (int)n) || (d != 0 && il <= 0);
            __syncwarp();
            if (lane == 0) { s_failed = fl; s_i_last = il; if (done) s_done = 1; }
            __syncwarp();
            if (!done) rebuild_slots(cur ^ 1);
        }
        grow = best >= 0;
        cur ^= 1;
        PF_MARK(4);  // insertion / rebuild
        __syncthreads();
        PF_MARK(5);  // wait
#ifdef POMFRET_JOIN_PROF
        pf_iter++;
#endif
    }
#ifdef POMFRET_JOIN_PROF
    if (lane == 0 && (warp == 0 || warp == 1 || warp == nwarps - 1))
        printf("JOINPROF w %u d %u warp %u n %u sites %u iters %d total %lld grow %lld score %lld wait2 %lld best %lld phaseD %lld wait3 %lld\n", w, d, warp, n,
               n_sites, pf_iter, clock64() - pf_t0, pf[0], pf[1], pf[2], pf[3], pf[4], pf[5]);
#endif
    // ---- 2x2 table over the far-side strict reads, :3888-3893 and :3940-3951 ----
    {
        const uint32_t *sid = (d == 0 ? P.ids_right_strict : P.ids_left_strict) + first;
        const uint32_t ns = d == 0 ? S.n_right_strict : S.n_left_strict;
        for (uint32_t i = tid; i < ns; i += nthreads) {
            const uint32_t id = sid[i];
            const uint8_t ref = (uint8_t)P.rs_hp[first + id];
            const uint8_t q = tags[id];
            if ((ref == 0 || ref == 1) && (q == 0 || q == 1)) atomicAdd(&s_tbl[ref * 2 + q], 1);
        }
        __syncthreads();
        if (tid == 0) {
            for (int t = 0; t < 4; t++) S.table[d][t] = s_tbl[t];
            S.n_order[d] = n_order;
        }
    }
}

#undef s_sid
#undef s_sseq

}  // namespace pomfret_gpu
#endif
