// intersection_vs_union all-pairs kernel (sm_100a).
//
// Replaces the pair loop of ComparableData.gen_comparable for score_func ==
// "intersection_vs_union" (/root/reference/napkon_string_matching/types/comparable_data.py:223-243,
// compare_terms :248-265, compare/score_functions.py:6-13).
//
// Data (pack.py): every level set carries a summary: `head` = exact bitset of its 64 most
// frequent vocabulary ids, `tail` = 64-bit signature of the rest, and its sizes.  The summaries
// are also laid out by compare_terms' step ("slot" t-1 holds level min(t, K-1) of each item),
// slot-major, so that the data of a block of consecutive items is contiguous per step.
//
// Work decomposition: a unit is one block of JT_THREADS right items x J_UNIT_LEFT left items.
// The right block's slots are staged once per unit in shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier; rows are padded so a block is one aligned run per step), one column per thread
// (lane i reads word i: conflict-free), together with one 128-bit "any" word per left item; the
// left items' slots are read through L1 where a pair needs them.  Each warp then runs a
// three-stage funnel over its 32 right items x the unit's left items with no block-wide barrier,
// re-compacting the survivors of every stage through a per-warp shared-memory queue (ballot +
// popc) so that each stage runs with 32 busy lanes:
//   A  ANY     (per item pair, ~8 integer ops) the OR of the summaries over the first D steps
//              (2^-D < threshold: later steps cannot lift a zero score over the threshold) shares
//              no bit => no level pair that matters shares a token => pair proven < threshold.
//   B  BOUND   (per step, integer + fp32 round-up, branch-free) exact head intersection by
//              popcount plus an upper bound of the tail intersection from the signatures bounds
//              every level score; compare_terms' weights are accumulated with directed rounding.
//              Six steps, straight-line, in two halves (steps 1..D, then the rest for the pairs
//              whose bound can still reach the threshold); later steps' weight is granted in full.
//   C  EXACT   per used level the exact |A & B| (popcount; only if both tail signatures collide,
//              a warp-cooperative intersection of the tail ids - one per pair, of the deepest
//              levels, when the levels are nested), the reference's int/int float64 division
//              and its accumulation order; score >= threshold in float64.
// Kept pairs go to a per-warp staging buffer and are flushed with one global atomic per ~50
// records as coalesced 16-byte stores.
#include "nsm_common.cuh"

namespace nsm {

constexpr int JT_THREADS = 128;   // right items per block (= threads per CTA)
constexpr int JT_CTAS = 5;        // resident CTAs per SM the kernel is sized for
constexpr int J_UNIT_LEFT = 512;  // left items per unit
constexpr int J_SLOTS = 10;       // steps staged per item (pack.py SLOT_CAP)
constexpr int J_RCP = 256;        // reciprocal table size
constexpr int J_WARPS = JT_THREADS / 32;
constexpr int J_OUT = 48;         // staged output records per warp
constexpr int J_UNROLL = 6;       // steps of stage B, all straight-line code
constexpr int J_AUNROLL = 4;      // left items per round of stage A
constexpr int J_QA = 32 + 32 * J_AUNROLL;  // queue A capacity
static_assert(J_UNIT_LEFT == 512 && JT_THREADS == 128, "queue entries are 9 + 7 bits; the tile table covers 512 rows");
// Stage A tiles: 32 right columns (a quarter of the block) x a range of left rows.  The warps of a
// CTA draw tiles from a shared counter; the ranges shrink towards the end of the unit (guided
// self-scheduling), so that the block barrier at the end waits for a 16-row tile at most.
constexpr int J_TILE_RANGES = 10;
__device__ const uint16_t g_tile_begin[J_TILE_RANGES + 1] = {0, 128, 256, 320, 384, 416, 448, 464, 480, 496, 512};
#ifndef NSM_J_DYN
#define NSM_J_DYN 1   // 1: the warps of a CTA draw stage A tiles from a shared counter; 0: fixed right quarters
#endif
#ifndef NSM_J_DYN_UNITS
#define NSM_J_DYN_UNITS 1  // 1: the CTAs draw their units from a device counter; 0: fixed stride
#endif
#ifndef NSM_J_COARSE
#define NSM_J_COARSE 1  // 1: coarse first half of stage B at low thresholds (see SPLIT == 0)
#endif
#ifndef NSM_J_CFLAT
#define NSM_J_CFLAT 1  // 1: branch-free step arithmetic in stage C's fast path
#endif
constexpr uint32_t NO_TWO = 0xffffffffu;   // JaccardParams::two_small: the >= 2 bits test is off

struct JaccardParams {
    nsm_sets_t L, R;
    nsm_job_t job;
    float thr_lo;        // filter threshold (see filter_threshold); -inf: everything passes
    uint32_t any_depth;  // D of stage A; 0: use the packed all-level item_any
    uint32_t bound_split;  // stage B tests the bound after this many steps (1..J_UNROLL)
    uint32_t two_small;    // TWO kernels: a step-1 level of at most this many ids makes its item "wild"
    uint32_t *unit_counter;  // device counter the CTAs draw their units from (zeroed before the launch)
    uint32_t dyn_tiles;    // 1: the warps of a CTA draw stage A tiles from a shared counter (see g_tile_begin)
    uint32_t n_lchunks, n_rblocks;
};

struct __align__(16) JaccardSmem {
    ulonglong2 r_ht[J_SLOTS][JT_THREADS];  // (head, tail) per step of my right block
    ulonglong2 l_any[J_UNIT_LEFT];         // stage A word of every left item of the unit
    union {
        nsm_pair_t pairs[J_OUT];  // NSM_OUT_PAIRS: staged records of the warp
        nsm_packet_t packet;      // NSM_OUT_PACKETS: the packet the warp is filling
        nsm_cpacket_t cpacket;    // NSM_OUT_CODED
    } out[J_WARPS];
    uint32_t r_info[J_SLOTS][JT_THREADS];  // size | fold count << 16
    uint32_t r_k[JT_THREADS];
    float rcp_up[J_RCP];
    ulonglong2 r_any[JT_THREADS];          // stage A word of every right column of the block
    float qm_ub[J_WARPS][64];   // partial bound of the entries of qm
    // the queues hold unit-local pair positions: left row (9 bits) << 7 | right column (7 bits)
    uint16_t qa[J_WARPS][J_QA];
    uint16_t qm[J_WARPS][64];   // survivors of the first half of stage B
    uint16_t qb[J_WARPS][64];
    uint16_t l_k[J_UNIT_LEFT];
    unsigned long long stats[NSM_N_STATS];
    uint64_t bar;  // mbarrier of the right block's bulk copies
    uint32_t tile_next;  // next stage A tile of the unit (the warps draw from it)
    uint32_t next_unit;  // the unit this CTA runs after the current one
};

// len(A & B) / len(A | B) as the correctly rounded float64 quotient of two small integers without
// the general division sequence: r = RN(1 / u) from a table, q0 = RN(i * r), e = RN(i - q0 * u)
// (one FMA), q = RN(q0 + e * r) (one FMA).  q equals RN(i / u) for every 0 <= i <= u <= 255 —
// checked exhaustively with exact rational arithmetic (tests/test_oracle.py); larger unions
// take __ddiv_rn.
struct RcpTable {
    double v[256];
    constexpr RcpTable() : v() {
        for (int u = 1; u < 256; ++u) v[u] = 1.0 / (double)u;
    }
};
__device__ const RcpTable g_rcp = RcpTable();

__device__ __forceinline__ double div_counts(uint32_t i, uint32_t u) {
    const double a = (double)i, b = (double)u;
    if (u < 256u) {
        const double r = g_rcp.v[u];
        const double q0 = __dmul_rn(a, r);
        return __fma_rn(__fma_rn(-q0, b, a), r, q0);
    }
    return __ddiv_rn(a, b);
}

__device__ __forceinline__ float pow2_neg(uint32_t k) {  // 2^-k, 0 when it underflows fp32
    return k <= 126 ? __int_as_float((127 - (int)k) << 23) : 0.0f;
}

// |A & B| of two sorted id lists, computed by the whole warp (all arguments warp-uniform).
// Small B (the usual case: a handful of tail ids): B lives in the lanes' registers and is
// broadcast by shuffle against 32 ids of A at a time, so the only memory latency is one load of
// each list.  Large B: every lane binary-searches its ids of A in B.
__device__ __forceinline__ uint32_t warp_intersect_count(const uint32_t *__restrict__ a, uint32_t na,
                                                         const uint32_t *__restrict__ b, uint32_t nb) {
    if (na == 0 || nb == 0) return 0;
    if (na < nb) {
        const uint32_t *tp = a; a = b; b = tp;
        const uint32_t tn = na; na = nb; nb = tn;
    }
    const unsigned lane = lane_id();
    uint32_t c = 0;
    if (nb <= 32) {
        const uint32_t yb = lane < nb ? __ldg(b + lane) : 0xffffffffu;
        for (uint32_t base = 0; base < na; base += 32) {
            const bool valid = base + lane < na;
            const uint32_t x = valid ? __ldg(a + base + lane) : 0u;
            bool hit = false;
            for (uint32_t j = 0; j < nb; ++j) hit |= (x == __shfl_sync(FULL_MASK, yb, j));
            c += __popc(__ballot_sync(FULL_MASK, hit && valid));
        }
    } else {
        for (uint32_t base = 0; base < na; base += 32) {
            const bool valid = base + lane < na;
            bool hit = false;
            if (valid) {
                const uint32_t x = __ldg(a + base + lane);
                uint32_t lo = 0, hi = nb;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (__ldg(b + mid) < x) lo = mid + 1; else hi = mid;
                }
                hit = lo < nb && __ldg(b + lo) == x;
            }
            c += __popc(__ballot_sync(FULL_MASK, hit));
        }
    }
    return c;
}

// Nested levels: one intersection of the two DEEPEST tail id lists gives the tail intersection of
// every step.  An id held from level ea on the left and eb on the right is shared from step
// max(1, ea, eb) on (step t uses level min(t, K-1) >= the entry level exactly when t >= it).
// Returns, packed 8 bits per step (t = 1..16), the number of shared tail ids at each step;
// all arguments warp-uniform, at most 255 shared ids (the caller checks min(na, nb) <= 255).
__device__ __forceinline__ ulonglong2 warp_intersect_steps(const uint32_t *__restrict__ a,
                                                           const uint8_t *__restrict__ ea, uint32_t na,
                                                           const uint32_t *__restrict__ b,
                                                           const uint8_t *__restrict__ eb, uint32_t nb,
                                                           uint32_t kmax) {
    ulonglong2 cnt = make_ulonglong2(0, 0);
    if (na == 0 || nb == 0) return cnt;
    if (na < nb) {
        const uint32_t *tp = a; a = b; b = tp;
        const uint8_t *te = ea; ea = eb; eb = te;
        const uint32_t tn = na; na = nb; nb = tn;
    }
    const unsigned lane = lane_id();
    const uint64_t ones = 0x0101010101010101ull;
    for (uint32_t base = 0; base < na; base += 32) {
        const bool valid = base + lane < na;
        const uint32_t x = valid ? __ldg(a + base + lane) : 0u;
        const uint32_t ex = valid ? (uint32_t)__ldg(ea + base + lane) : 0u;
        uint32_t d = 0xffffu;  // first step at which my id is shared; 0xffff: never
        if (nb <= 32) {
            const uint32_t yb = lane < nb ? __ldg(b + lane) : 0xffffffffu;
            const uint32_t ey = lane < nb ? (uint32_t)__ldg(eb + lane) : 0u;
            for (uint32_t j = 0; j < nb; ++j) {
                const uint32_t y = __shfl_sync(FULL_MASK, yb, j), e = __shfl_sync(FULL_MASK, ey, j);
                if (valid && x == y) d = max(max(ex, e), 1u);
            }
        } else if (valid) {
            uint32_t lo = 0, hi = nb;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(b + mid) < x) lo = mid + 1; else hi = mid;
            }
            if (lo < nb && __ldg(b + lo) == x) d = max(max(ex, (uint32_t)__ldg(eb + lo)), 1u);
        }
        // my id counts for every step t >= d: a one in the bytes d-1 .. 15, summed over the warp
        // (at most 32 per byte and chunk, at most 255 in total: bytes never carry into each other)
        const uint64_t lo = d <= 8 ? ones << (8 * (d - 1)) : 0ull;
        const uint64_t hi = d <= 8 ? ones : (d <= 16 ? ones << (8 * (d - 9)) : 0ull);
        cnt.x += (uint64_t)__reduce_add_sync(FULL_MASK, (uint32_t)lo) |
                 ((uint64_t)__reduce_add_sync(FULL_MASK, (uint32_t)(lo >> 32)) << 32);
        cnt.y += (uint64_t)__reduce_add_sync(FULL_MASK, (uint32_t)hi) |
                 ((uint64_t)__reduce_add_sync(FULL_MASK, (uint32_t)(hi >> 32)) << 32);
    }
    (void)kmax;
    return cnt;
}

// Upper bound of |A & B| of one level pair from the summaries; exact when the tails share no
// bit.  Shared tail bits + the ids either side folded onto an occupied bit bound the shared tail
// ids (a fold count of 255 is saturated: fall back to min(|A|, |B|)).  Branch-free.
__device__ __forceinline__ uint32_t bound_intersection(const ulonglong2 &A, uint32_t ia,
                                                       const ulonglong2 &B, uint32_t ib,
                                                       bool exact_bits) {
    const uint64_t tb = A.y & B.y;
    const uint32_t ex = min(ia >> 16, ib >> 16);
    uint32_t it = __popcll(tb) + (exact_bits ? 0u : ex + (ex == 255u ? 0x10000u : 0u));
    it = tb ? it : 0u;
    return min(__popcll(A.x & B.x) + it, min(ia & 0xffffu, ib & 0xffffu));
}

// Stage A word of the TWO kernels (depth D = 1, threshold so high that ONE shared token cannot
// lift a pair over it): x = head bitset of the step-1 level, y = its tail signature folded to 62
// bits (bits 62, 63 onto 60, 61) plus two flag bits.  An item is "wild" when shared signature
// bits may undercount shared ids (two of its tail ids share a bit) or when its level is so small
// that one shared id could be enough; a pair with a wild side is tested for >= 1 shared bit as
// before, every other pair for >= 2 shared bits.  Left words carry (wild << 63 | 1 << 62), right
// words (1 << 63 | wild << 62), so that bits 63 / 62 of the AND are the two wild flags.
__device__ __forceinline__ ulonglong2 two_word(ulonglong2 ht, uint32_t info, uint32_t small, bool left) {
    const uint64_t low = ht.y & 0x3fffffffffffffffull, top = (ht.y >> 62) << 60;
    const bool wild = (info >> 16) != 0 || (low & top) != 0 || (info & 0xffffu) <= small;
    const uint64_t w = wild ? 1ull : 0ull;
    return make_ulonglong2(ht.x, low | top | (left ? (w << 63) | (1ull << 62) : (1ull << 63) | (w << 62)));
}

template <bool DEEP, int SPLIT, bool TWO>
__global__ void __launch_bounds__(JT_THREADS, JT_CTAS)
jaccard_allpairs_kernel(const JaccardParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    JaccardSmem &s = *reinterpret_cast<JaccardSmem *>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long *count = reinterpret_cast<unsigned long long *>(p.job.out_count);
    const bool flat = p.job.flat != 0;
    const bool exact_bits = p.L.exact_bits != 0 && p.R.exact_bits != 0;
    const bool nested_mode = !exact_bits && p.L.nested != 0 && p.R.nested != 0;
    const double thr = p.job.threshold;
    const bool pass_all = !(p.thr_lo > -INFINITY) && !(p.thr_lo != p.thr_lo);  // thr <= 0
    const uint32_t SL = p.L.n_slots, SR = p.R.n_slots;
    const ulonglong2 *l_slot_ht = reinterpret_cast<const ulonglong2 *>(p.L.slot_ht);

    for (unsigned u = tid; u < J_RCP; u += JT_THREADS) s.rcp_up[u] = u ? __frcp_ru((float)u) : 0.0f;
    if (tid < NSM_N_STATS) s.stats[tid] = 0;
    if (tid == 0) mbar_init(&s.bar, 1);
    uint32_t bar_parity = 0;
    unsigned long long st_cand = 0, st_evals = 0, st_merges = 0, st_bound = 0, st_kept = 0;
    uint32_t out_n = 0;  // warp-uniform fill of s.out[warp]
    const bool packets = p.job.out_mode == NSM_OUT_PACKETS, coded = p.job.out_mode == NSM_OUT_CODED;

    auto flush_out = [&]() {
        if (out_n == 0) return;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(count, (unsigned long long)out_n);
        base = __shfl_sync(FULL_MASK, base, 0);
        const double2 *src = reinterpret_cast<const double2 *>(&s.out[warp].pairs[0]);
        for (uint32_t i = lane; i < out_n; i += 32) {
            const unsigned long long pos = base + i;
            if (pos < p.job.out_capacity)
                reinterpret_cast<double2 *>(p.job.out_pairs)[pos] = src[i];
            else
                atomicOr(p.job.out_flags, NSM_FLAG_OVERFLOW);
        }
        __syncwarp();
        out_n = 0;
    };
    // NSM_OUT_PACKETS: the warp's packet goes out as 31 16-byte stores, one atomic per packet.
    // Packets are filled to the last slot; only the one a warp holds at the end of a unit is partial.
    auto flush_packet = [&](uint32_t n_rec, uint32_t l0, uint32_t r0) {
        nsm_packet_t &pk = s.out[warp].packet;
        unsigned long long pos = 0;
        if (lane == 0) {
            pk.left0 = l0; pk.right0 = r0; pk.count = n_rec; pk.reserved_ = 0;
            pos = atomicAdd(count, 1ull);
        }
        __syncwarp();  // the header (and the callers' record stores) before the vector reads below
        pos = __shfl_sync(FULL_MASK, pos, 0);
        if (pos < p.job.out_capacity) {
            constexpr uint32_t VECS = sizeof(nsm_packet_t) / 16;
            static_assert(sizeof(nsm_packet_t) == 496 && VECS <= 32, "one 16-byte vector per lane");
            if (lane < VECS)
                reinterpret_cast<uint4 *>(reinterpret_cast<nsm_packet_t *>(p.job.out_pairs) + pos)[lane] =
                    reinterpret_cast<const uint4 *>(&pk)[lane];
        } else if (lane == 0) {
            atomicOr(p.job.out_flags, NSM_FLAG_OVERFLOW);
        }
        __syncwarp();
    };

    // NSM_OUT_CODED: 16 16-byte stores per packet
    auto flush_cpacket = [&](uint32_t n_rec, uint32_t l0, uint32_t r0) {
        nsm_cpacket_t &pk = s.out[warp].cpacket;
        unsigned long long pos = 0;
        if (lane == 0) {
            pk.left0 = l0; pk.right0 = r0; pk.count = n_rec; pk.reserved_ = 0;
            pos = atomicAdd(count, 1ull);
        }
        __syncwarp();
        pos = __shfl_sync(FULL_MASK, pos, 0);
        if (pos < p.job.out_capacity) {
            constexpr uint32_t VECS = sizeof(nsm_cpacket_t) / 16;
            static_assert(sizeof(nsm_cpacket_t) == 256, "16 vectors");
            if (lane < VECS)
                reinterpret_cast<uint4 *>(reinterpret_cast<nsm_cpacket_t *>(p.job.out_pairs) + pos)[lane] =
                    reinterpret_cast<const uint4 *>(&pk)[lane];
        } else if (lane == 0) {
            atomicOr(p.job.out_flags, NSM_FLAG_OVERFLOW);
        }
        __syncwarp();
    };
    // Slot of a score in the dictionary (its 16-bit code), inserting it when a slot within reach
    // of its hash is free; 0xffffffff when there is none (the pair then goes out uncoded).
    // Keys are only ever added, so a slot that shows the key — even from a stale cache line — is right.
    auto dict_code = [&](double score) -> uint32_t {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(score);
        unsigned long long *table = reinterpret_cast<unsigned long long *>(p.job.out_dict);
        const uint32_t h = (uint32_t)((bits * 0x9E3779B97F4A7C15ull) >> 48);
#pragma unroll 1
        for (uint32_t probe = 0; probe < 8; ++probe) {
            const uint32_t slot = (h + probe) & (NSM_DICT_SLOTS - 1u);
            unsigned long long k = __ldcg(table + slot);
            if (k == NSM_DICT_FREE) k = atomicCAS(table + slot, NSM_DICT_FREE, bits), k = k == NSM_DICT_FREE ? bits : k;
            if (k == bits) return slot;
        }
        return 0xffffffffu;
    };

    // summary of one item at step t >= 1: from the slot arrays, or from the CSR arrays when the
    // item has more levels than slots and t lies beyond them
    auto left_level = [&](uint32_t item, uint32_t t, uint32_t k, uint32_t &info) {
        if (DEEP && t > SL && k > SL + 1) {
            const uint32_t g = __ldg(p.L.item_level_off + item) + min(t, k - 1);
            info = __ldg(p.L.level_info + g);
            return make_ulonglong2(__ldg(p.L.level_head + g), __ldg(p.L.level_tail + g));
        }
        const uint32_t at = (min(t, SL) - 1) * p.L.slot_stride + item;   // < 2^32: checked by the host
        info = __ldg(p.L.slot_info + at);
        return __ldg(l_slot_ht + at);
    };
    auto right_level = [&](uint32_t rc, uint32_t item, uint32_t t, uint32_t k, uint32_t &info) {
        if (DEEP && t > SR && k > SR + 1) {
            const uint32_t g = __ldg(p.R.item_level_off + item) + min(t, k - 1);
            info = __ldg(p.R.level_info + g);
            return make_ulonglong2(__ldg(p.R.level_head + g), __ldg(p.R.level_tail + g));
        }
        const uint32_t sr = min(t, SR) - 1;
        info = s.r_info[sr][rc];
        return s.r_ht[sr][rc];
    };

    // Units are handed out through a device counter: a CTA takes its next unit when it is done with
    // the current one (the first one is its block index), so no CTA's share depends on how the cost
    // of the units varies with their position.  Thread 0 draws the next unit while the current one
    // is staged; everybody picks it up behind the staging barrier.
    const uint32_t n_units = p.n_lchunks * p.n_rblocks;
    uint32_t unit = blockIdx.x;
    while (unit < n_units) {
        const uint32_t rb = unit / p.n_lchunks, lchunk = unit - rb * p.n_lchunks;
        const uint32_t r0 = rb * JT_THREADS;
        const uint32_t l0 = p.job.l_row_begin + lchunk * J_UNIT_LEFT;
        const uint32_t nl = min((uint32_t)J_UNIT_LEFT, p.job.l_row_end - l0);

        __syncthreads();  // previous unit fully consumed
        // ---- stage the right block's slots: 2 x SR TMA bulk copies, one column per thread ----
        // (slot rows are padded to 128 items, so every block is a full, 16-byte aligned run)
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(&s.bar, SR * (JT_THREADS * 16u + JT_THREADS * 4u));
            for (uint32_t sl = 0; sl < SR; ++sl) {
                const size_t at = (size_t)sl * p.R.slot_stride + r0;
                bulk_copy_g2s(&s.r_ht[sl][0], reinterpret_cast<const ulonglong2 *>(p.R.slot_ht) + at,
                              JT_THREADS * 16u, &s.bar);
                bulk_copy_g2s(&s.r_info[sl][0], p.R.slot_info + at, JT_THREADS * 4u, &s.bar);
            }
        }
        const uint32_t r = r0 + tid;
        s.r_k[tid] = r < p.R.n_items ? __ldg(p.R.item_k + r) : 0u;
        if (tid == 0) {
            s.tile_next = 0;
            s.next_unit = p.unit_counter ? gridDim.x + atomicAdd(p.unit_counter, 1u) : unit + gridDim.x;
        }
        // ---- the unit's left items: stage A word and level count (rows beyond nl: a word that
        // ---- shares no bit with anything, so stage A needs no bounds test) ----------------
        for (uint32_t li = tid; li < (uint32_t)J_UNIT_LEFT; li += JT_THREADS) {
            ulonglong2 any = make_ulonglong2(0, 0);
            uint32_t k = 0;
            if (li < nl) {
                k = __ldg(p.L.item_k + l0 + li);
                if (pass_all || k == 0) {
                    any.x = any.y = ~0ull;
                } else if (p.any_depth == 0) {
                    any = __ldg(reinterpret_cast<const ulonglong2 *>(p.L.item_any) + l0 + li);
                } else if (TWO) {
                    any = two_word(__ldg(l_slot_ht + l0 + li), __ldg(p.L.slot_info + l0 + li), p.two_small, true);
                } else {
                    for (uint32_t sl = 0; sl < min(p.any_depth, SL); ++sl) {
                        const ulonglong2 ht = __ldg(l_slot_ht + (size_t)sl * p.L.slot_stride + l0 + li);
                        any.x |= ht.x; any.y |= ht.y;
                    }
                }
            }
            s.l_any[li] = any;
            s.l_k[li] = (uint16_t)min(k, 0xffffu);
        }
        mbar_wait(&s.bar, bar_parity);  // the right block has landed
        bar_parity ^= 1u;
        {   // stage A word of my right column.  A column beyond the cohort shares no bit with
            // anything; threshold <= 0 keeps every pair; an item without levels must reach stage C,
            // which tells "both empty: score 0" from "one empty: IndexError upstream"
            ulonglong2 any = make_ulonglong2(0, 0);
            if (r < p.R.n_items) {
                if (pass_all || s.r_k[tid] == 0) {
                    any.x = any.y = ~0ull;
                } else if (TWO) {
                    any = two_word(s.r_ht[0][tid], s.r_info[0][tid], p.two_small, false);
                } else if (p.any_depth == 0) {
                    any = __ldg(reinterpret_cast<const ulonglong2 *>(p.R.item_any) + r);
                } else {
                    for (uint32_t sl = 0; sl < min(p.any_depth, SR); ++sl) {
                        const ulonglong2 ht = s.r_ht[sl][tid];
                        any.x |= ht.x; any.y |= ht.y;
                    }
                }
            }
            s.r_any[tid] = any;
        }
        __syncthreads();
        const uint32_t unit_after = s.next_unit;

        // Stage A runs tile by tile (see g_tile_begin): a warp whose columns hold frequent tokens
        // (many survivors, long stages B and C) simply draws fewer tiles.
        const uint32_t n_tiles = 4u * J_TILE_RANGES;
        uint32_t qa_n = 0, qm_n = 0, qb_n = 0;  // warp-uniform
        uint32_t li_next = 0, li_end = 0, rc_mine = (warp << 5) | lane;  // the tile the warp is in
        uint64_t rany_h = 0, rany_t = 0;
        bool unit_done = false;
        if (!p.dyn_tiles) {   // fixed quarters: one tile per warp
            const ulonglong2 w = s.r_any[rc_mine];
            rany_h = w.x; rany_t = w.y; li_end = (nl + 3u) & ~3u;
        }
        while (true) {
            // ---- stage A: left items x my right item ----------------------------------------
            while (qa_n < 32 && !unit_done) {
                if (li_next >= li_end) {
                    if (!p.dyn_tiles) { unit_done = true; break; }
                    uint32_t tile = 0;
                    if (lane == 0) tile = atomicAdd(&s.tile_next, 1u);
                    tile = __shfl_sync(FULL_MASK, tile, 0);
                    if (tile >= n_tiles) { unit_done = true; break; }
                    rc_mine = ((tile & 3u) << 5) | lane;
                    li_next = g_tile_begin[tile >> 2];
                    li_end = min((uint32_t)g_tile_begin[(tile >> 2) + 1], (nl + 3u) & ~3u);
                    const ulonglong2 w = s.r_any[rc_mine];
                    rany_h = w.x; rany_t = w.y;
                    continue;   // a range beyond the unit's rows is empty
                }
                // four left items per round: four independent load -> test -> vote chains
                bool pass[J_AUNROLL];
                unsigned m[J_AUNROLL];
#pragma unroll
                for (int u = 0; u < J_AUNROLL; ++u) {
                    const ulonglong2 lany = s.l_any[li_next + u];
                    if (TWO) {
                        // shared bits of the step-1 levels: at least two, or one next to a wild flag
                        const uint64_t z = lany.x & rany_h, w = lany.y & rany_t;
                        const uint32_t z0 = (uint32_t)z, z1 = (uint32_t)(z >> 32), w0 = (uint32_t)w,
                                       w1 = (uint32_t)(w >> 32), wc = w1 & 0x3fffffffu;
                        const uint32_t zz = z0 | z1, ww = w0 | wc, any1 = zz | ww;
                        const uint32_t multi = (any1 & (any1 - 1u)) | (z0 & z1) | (w0 & wc) | (zz & ww);
                        pass[u] = any1 != 0 && (multi | (w1 >> 30)) != 0;
                    } else {
                        pass[u] = ((lany.x & rany_h) | (lany.y & rany_t)) != 0;
                    }
                    m[u] = __ballot_sync(FULL_MASK, pass[u]);
                }
#pragma unroll
                for (int u = 0; u < J_AUNROLL; ++u) {
                    if (pass[u])
                        s.qa[warp][qa_n + __popc(m[u] & lanemask_lt())] = ((li_next + u) << 7) | rc_mine;
                    qa_n += __popc(m[u]);
                }
                li_next += J_AUNROLL;
            }
            if (qa_n == 0 && qm_n == 0 && qb_n == 0 && unit_done) break;
            __syncwarp();

            // ---- stage B: fp32 upper bound of one surviving pair per lane -----------------
            // Two halves with a queue in between: steps 1..D first (D as in stage A: once they
            // are known, bound + remaining weight can already fall short of the threshold), then
            // steps D+1..J_UNROLL for the pairs that are still alive.
            auto bound_step = [&](uint32_t t, uint32_t li, uint32_t rc, uint32_t kmax, float ub) {
                // steps beyond kmax read a repeated level and get weight 0
                const uint32_t sr = min(t, SR) - 1;
                const uint32_t at = (min(t, SL) - 1) * p.L.slot_stride + l0 + li;
                const uint32_t ia = __ldg(p.L.slot_info + at), ib = s.r_info[sr][rc];
                const uint32_t ih = bound_intersection(__ldg(l_slot_ht + at), ia, s.r_ht[sr][rc], ib,
                                                       exact_bits);
                const uint32_t uh = min((ia & 0xffffu) + (ib & 0xffffu) - ih,
                                        (uint32_t)J_RCP - 1);  // 1/255 >= 1/u beyond
                const float w = t <= kmax ? (flat ? 1.0f : pow2_neg(t)) : 0.0f;
                return __fmaf_ru(__fmul_ru((float)ih, s.rcp_up[uh]), w, ub);
            };
            if (qa_n >= 32 || (unit_done && qa_n)) {
                const uint32_t take = min(qa_n, 32u);
                qa_n -= take;
                const bool active = lane < take;
                const uint32_t entry = active ? s.qa[warp][qa_n + lane] : 0u;
                const uint32_t li = entry >> 7, rc = entry & 127u;
                const uint32_t kl = s.l_k[li], c_kr = s.r_k[rc];
                bool pass = active;
                // the category predicate runs here, once per round of survivors, not per pair of
                // stage A (it is off in the shipped configuration)
                if (p.job.cat_mode)
                    pass = pass && keep_categories(p.job.cat_mode, __ldg(p.job.l_cat + l0 + li),
                                                   __ldg(p.job.r_cat + min(r0 + rc, p.R.n_items - 1)));
                float ub = 0.0f;
                if (pass && !pass_all && kl != 0 && c_kr != 0) {
                    const uint32_t kmax = flat ? 1u : max(kl, c_kr);
                    ++st_bound;
                    if (SPLIT == 0) {
                        // COARSE first half (low thresholds, D >= 3; nested levels on both sides, every
                        // level in a slot).  A pair that shares no bit at step 2 has J_1 = J_2 = 0
                        // (levels are nested: step 1's sets lie inside step 2's).  For its steps
                        // 3..T (T = J_UNROLL): I_t <= I_T <= ih, the bound from step T's summaries, and
                        // the unions only grow, U_t >= U_3 >= a_3 + b_3 - min(ih, a_3, b_3) = umin, so
                        // every J_t <= jc = min(1, ih / umin) and
                        //   score <= (2^-2 - 2^-min(T, kmax)) jc + [kmax > T] (2^-T - 2^-kmax).
                        // Pairs that share a bit at step 2, or whose coarse bound reaches the
                        // threshold, get the per-step bound of all T steps in the second half.
                        const uint32_t sl2 = min(2u, SL) - 1, sl3 = min(3u, SL) - 1, slT = min((uint32_t)J_UNROLL, SL) - 1;
                        const uint32_t sr2 = min(2u, SR) - 1, sr3 = min(3u, SR) - 1, srT = min((uint32_t)J_UNROLL, SR) - 1;
                        const uint32_t item = l0 + li;
                        const ulonglong2 A2 = __ldg(l_slot_ht + (sl2 * p.L.slot_stride + item));
                        const ulonglong2 AT = __ldg(l_slot_ht + (slT * p.L.slot_stride + item));
                        const uint32_t iaT = __ldg(p.L.slot_info + (slT * p.L.slot_stride + item));
                        const uint32_t ia3 = __ldg(p.L.slot_info + (sl3 * p.L.slot_stride + item));
                        const ulonglong2 B2 = s.r_ht[sr2][rc];
                        const bool share2 = ((A2.x & B2.x) | (A2.y & B2.y)) != 0;
                        const uint32_t ih = bound_intersection(AT, iaT, s.r_ht[srT][rc], s.r_info[srT][rc], exact_bits);
                        const uint32_t a3 = ia3 & 0xffffu, b3 = s.r_info[sr3][rc] & 0xffffu;
                        const uint32_t umin = a3 + b3 - min(ih, min(a3, b3));
                        // umin == 0: both step-3 sets empty (0 / 0 upstream): let stage C see the pair
                        const float jc = umin == 0 ? 1.0f
                            : fminf(1.0f, __fmul_ru((float)ih, s.rcp_up[min(umin, (uint32_t)J_RCP - 1)]));
                        const uint32_t kt = min(kmax, (uint32_t)J_UNROLL);
                        const float wsum = kt > 2 ? __fsub_ru(0.25f, pow2_neg(kt)) : 0.0f;
                        const float grant = kmax > (uint32_t)J_UNROLL
                            ? __fsub_ru(pow2_neg(J_UNROLL), pow2_neg(kmax)) : 0.0f;
                        pass = share2 || __fmaf_ru(jc, wsum, grant) >= p.thr_lo;
                    } else {
#pragma unroll
                        for (uint32_t t = 1; t <= (uint32_t)J_UNROLL; ++t)
                            if (t <= (uint32_t)SPLIT) ub = bound_step(t, li, rc, kmax, ub);
                        // weights still to come after step D: 2^-D - 2^-kmax
                        const float rem = kmax > (uint32_t)SPLIT
                            ? __fsub_ru(pow2_neg(SPLIT), pow2_neg(kmax)) : 0.0f;
                        pass = __fadd_ru(ub, rem) >= p.thr_lo;
                    }
                }
                const unsigned m = __ballot_sync(FULL_MASK, pass);
                if (pass) {
                    const uint32_t at = qm_n + __popc(m & lanemask_lt());
                    s.qm[warp][at] = entry;
                    s.qm_ub[warp][at] = ub;
                }
                qm_n += __popc(m);
                __syncwarp();
            }
            if (qm_n >= 32 || (unit_done && qa_n == 0 && qm_n)) {
                const uint32_t take = min(qm_n, 32u);
                qm_n -= take;
                const bool active = lane < take;
                const uint32_t entry = active ? s.qm[warp][qm_n + lane] : 0u;
                float ub = active ? s.qm_ub[warp][qm_n + lane] : 0.0f;
                const uint32_t li = entry >> 7, rc = entry & 127u;
                const uint32_t kl = s.l_k[li], c_kr = s.r_k[rc];
                bool pass = active;
                if (active && !pass_all && kl != 0 && c_kr != 0) {
                    const uint32_t kmax = flat ? 1u : max(kl, c_kr);
#pragma unroll
                    for (uint32_t t = 1; t <= (uint32_t)J_UNROLL; ++t)
                        if (t > (uint32_t)SPLIT) ub = bound_step(t, li, rc, kmax, ub);
                    // steps beyond the unrolled ones are not bounded individually: all their
                    // weight, 2^-UNROLL - 2^-kmax, is granted (pairs this lets through are within
                    // 2^-UNROLL of the threshold and get their exact score in stage C)
                    if (kmax > (uint32_t)J_UNROLL)
                        ub = __fadd_ru(ub, __fsub_ru(pow2_neg(J_UNROLL), pow2_neg(kmax)));
                    pass = ub >= p.thr_lo;
                }
                const unsigned m = __ballot_sync(FULL_MASK, pass);
                if (pass) s.qb[warp][qb_n + __popc(m & lanemask_lt())] = entry;
                qb_n += __popc(m);
                __syncwarp();
            }

            // ---- stage C: exact score of one candidate per lane ---------------------------
            // The step loop is warp-uniform (lanes whose pair has fewer steps idle) so that the
            // rare tail intersections can be computed by the whole warp.
            if (qb_n >= 32 || (unit_done && qa_n == 0 && qm_n == 0 && qb_n)) {
                const uint32_t take = min(qb_n, 32u);
                qb_n -= take;
                const bool active = lane < take;
                const uint32_t entry = active ? s.qb[warp][qb_n + lane] : 0u;
                const uint32_t li = entry >> 7, rc = entry & 127u;
                const uint32_t c_l = l0 + li, c_r = r0 + rc;
                const uint32_t kl = s.l_k[li], c_kr = s.r_k[rc];
                bool ok = active;
                uint32_t kmax = 0;
                if (active) {
                    ++st_cand;
                    if (kl == 0 || c_kr == 0) {
                        // both empty: compare_terms returns 0; one empty: IndexError upstream
                        if (kl != c_kr) { atomicOr(p.job.out_flags, NSM_FLAG_EMPTY_ITEM); ok = false; }
                    } else {
                        kmax = flat ? 1u : max(kl, c_kr);
                    }
                }
                // first level index of both items: only the tail-collision path needs them, but
                // fetching them now takes their latency off that path
                const uint32_t lg0 = exact_bits ? 0u : __ldg(p.L.item_level_off + c_l);
                const uint32_t rg0 = exact_bits ? 0u : __ldg(p.R.item_level_off + c_r);
                const uint32_t kmax_warp = __reduce_max_sync(FULL_MASK, kmax);
                double score = 0.0, w = flat ? 2.0 : 1.0;
                uint32_t pjl = 0xffffffffu, pjr = 0xffffffffu, inter = 0, uni = 1, a = 0, b = 0;
                // Nested levels (what gen_comp_value produces): the tail intersections of all steps
                // come from ONE cooperative intersection of the deepest levels' tail ids.
                bool nest_ok = nested_mode && kmax != 0 && kmax <= 16;
                ulonglong2 tail_steps = make_ulonglong2(0, 0);
                if (nested_mode) {  // kernel-uniform
                    bool need_deep = false;
                    uint32_t off_a = 0, off_b = 0, n_a = 0, n_b = 0;
                    if (nest_ok) {
                        uint32_t ia, ib;
                        const ulonglong2 A = left_level(c_l, kmax, kl, ia);
                        const ulonglong2 B = right_level(rc, c_r, kmax, c_kr, ib);
                        if (A.y & B.y) {
                            const uint32_t hl = __popcll(A.x), hr = __popcll(B.x);
                            n_a = (ia & 0xffffu) - hl; n_b = (ib & 0xffffu) - hr;
                            if (min(n_a, n_b) > 255u) {
                                nest_ok = false;  // counts would not fit: per-level path below
                            } else {
                                const uint32_t gl = lg0 + kl - 1, gr = rg0 + c_kr - 1;
                                const uint64_t t2l = __ldg(p.L.level_tail2 + gl), t2r = __ldg(p.R.level_tail2 + gr);
                                off_a = __ldg(p.L.level_tok_off + gl) + hl;
                                off_b = __ldg(p.R.level_tok_off + gr) + hr;
                                need_deep = (t2l & t2r) != 0;
                            }
                        }
                    }
                    unsigned todo = __ballot_sync(FULL_MASK, need_deep);
                    while (todo) {
                        const int src = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const uint32_t oa = __shfl_sync(FULL_MASK, off_a, src), ob = __shfl_sync(FULL_MASK, off_b, src);
                        const ulonglong2 c = warp_intersect_steps(
                            p.L.tok + oa, p.L.tok_entry + oa, __shfl_sync(FULL_MASK, n_a, src),
                            p.R.tok + ob, p.R.tok_entry + ob, __shfl_sync(FULL_MASK, n_b, src),
                            __shfl_sync(FULL_MASK, kmax, src));
                        if ((int)lane == src) { tail_steps = c; ++st_merges; }
                    }
                }
                const uint32_t kl1 = max(kl, 1u);
                // Fast path (warp-uniform): no lane needs a per-level token intersection inside
                // the step loop (exact tail bits, or the tail counts of all steps are already in
                // tail_steps).  The steps are then independent of one another up to the final
                // accumulation, so they are scored two at a time with all loads issued first and
                // the left summaries of the next two steps in flight; only the float64
                // accumulation runs in step order, as in the reference.
                const bool fast_warp = __all_sync(FULL_MASK, exact_bits || nest_ok || kmax == 0);
                if (fast_warp) {
                    bool zero_union = false;
                    ulonglong2 An[2];
                    uint32_t ian[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) An[q] = left_level(c_l, 1 + q, kl1, ian[q]);
                    for (uint32_t t0 = 1; t0 <= kmax_warp; t0 += 2) {
                        ulonglong2 A[2], B[2];
                        uint32_t ia[2], ib[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            A[q] = An[q]; ia[q] = ian[q];
                            B[q] = right_level(rc, c_r, t0 + q, c_kr, ib[q]);
                        }
                        if (t0 + 2 <= kmax_warp) {
#pragma unroll
                            for (int q = 0; q < 2; ++q) An[q] = left_level(c_l, t0 + 2 + q, kl1, ian[q]);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t t = t0 + q;
                            uint32_t it = __popcll(A[q].x & B[q].x);
                            const uint64_t tb = A[q].y & B[q].y;
#if NSM_J_CFLAT
                            // no data-dependent branch: a zero intersection adds +0.0 (0 / u = +0.0 and
                            // score + 0.0 * w == score), a step beyond the pair's schedule adds nothing
                            if (exact_bits) {
                                it += __popcll(tb);
                            } else {   // t is warp-uniform: one shift of the packed per-step counts
                                const uint32_t cnt = (uint32_t)((t <= 8 ? tail_steps.x >> (8 * (t - 1))
                                                                        : tail_steps.y >> (8 * ((t - 9) & 7))) & 0xffu);
                                it += (tb != 0 && t <= 16) ? cnt : 0u;
                            }
                            const uint32_t un = (ia[q] & 0xffffu) + (ib[q] & 0xffffu) - it;
                            const bool on = t <= kmax;
                            st_evals += on ? 1u : 0u;
                            w *= 0.5;
                            const double sc = __fma_rn(div_counts(it, un), w, score);
                            score = on ? sc : score;
                            zero_union |= on && un == 0;  // 0 / 0: ZeroDivisionError upstream
#else
                            if (exact_bits) {
                                it += __popcll(tb);
                            } else if (tb && t <= 16) {
                                it += (uint32_t)((t <= 8 ? tail_steps.x >> (8 * (t - 1))
                                                         : tail_steps.y >> (8 * (t - 9))) & 0xffu);
                            }
                            const uint32_t un = (ia[q] & 0xffffu) + (ib[q] & 0xffffu) - it;
                            if (t <= kmax) {
                                ++st_evals;
                                w *= 0.5;
                                if (it) score = __fma_rn(div_counts(it, un), w, score);
                                else zero_union |= un == 0;  // 0 / 0: ZeroDivisionError upstream
                            }
#endif
                        }
                    }
                    if (zero_union) {
                        atomicOr(p.job.out_flags, NSM_FLAG_ZERO_UNION);
                        ok = false;
                    }
                }
                // General path: the left summaries come through L2: keep the loads of the next
                // two steps in flight while the current step is scored
                uint32_t ia1 = 0, ia2 = 0;
                ulonglong2 A1 = make_ulonglong2(0, 0), A2 = A1;
                if (!fast_warp && kmax_warp >= 1) A1 = left_level(c_l, 1, kl1, ia1);
                if (!fast_warp && kmax_warp >= 2) A2 = left_level(c_l, 2, kl1, ia2);
                for (uint32_t t = 1; !fast_warp && t <= kmax_warp; ++t) {
                    const bool on = t <= kmax;
                    const uint32_t jl = flat ? 0u : min(t, kl - 1), jr = flat ? 0u : min(t, c_kr - 1);
                    bool need = false;
                    uint32_t hl = 0, hr = 0;
                    const ulonglong2 A = A1;
                    const uint32_t ia = ia1;
                    A1 = A2; ia1 = ia2;
                    if (t + 2 <= kmax_warp) A2 = left_level(c_l, t + 2, kl1, ia2);
                    if (on) {
                        ++st_evals;
                        if (jl != pjl || jr != pjr) {
                            pjl = jl; pjr = jr;
                            uint32_t ib;
                            const ulonglong2 B = right_level(rc, c_r, t, c_kr, ib);
                            a = ia & 0xffffu; b = ib & 0xffffu;
                            inter = __popcll(A.x & B.x);
                            const uint64_t tb = A.y & B.y;
                            if (tb) {
                                if (exact_bits) {
                                    inter += __popcll(tb);
                                } else if (nest_ok) {
                                    inter += (uint32_t)((t <= 8 ? tail_steps.x >> (8 * (t - 1))
                                                                : tail_steps.y >> (8 * (t - 9))) & 0xffu);
                                } else {  // ids are sorted: the tail ids follow the head ids
                                    need = true; hl = __popcll(A.x); hr = __popcll(B.x);
                                }
                            }
                            uni = a + b - inter;
                        }
                    }
                    if (__any_sync(FULL_MASK, need)) {
                        uint32_t off_a = 0, off_b = 0;
                        if (need) {
                            const uint32_t gl = lg0 + jl, gr = rg0 + jr;
                            // four independent loads, one round trip
                            const uint64_t t2l = __ldg(p.L.level_tail2 + gl), t2r = __ldg(p.R.level_tail2 + gr);
                            off_a = __ldg(p.L.level_tok_off + gl) + hl;
                            off_b = __ldg(p.R.level_tok_off + gr) + hr;
                            // a second, independent signature rules most collisions out
                            need = (t2l & t2r) != 0;
                        }
                        unsigned todo = __ballot_sync(FULL_MASK, need);
                        while (todo) {
                            const int src = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const uint32_t c = warp_intersect_count(
                                p.L.tok + __shfl_sync(FULL_MASK, off_a, src), __shfl_sync(FULL_MASK, a - hl, src),
                                p.R.tok + __shfl_sync(FULL_MASK, off_b, src), __shfl_sync(FULL_MASK, b - hr, src));
                            if ((int)lane == src) { inter += c; uni = a + b - inter; ++st_merges; }
                        }
                    }
                    if (on) {
                        w *= 0.5;
                        if (inter) {
                            // len(A & B) / len(A | B): int / int true division; score += s * w
                            score = __fma_rn(div_counts(inter, uni), w, score);
                        } else if (uni == 0) {  // 0 / 0: ZeroDivisionError upstream
                            atomicOr(p.job.out_flags, NSM_FLAG_ZERO_UNION);
                            ok = false;
                        }  // else 0 / uni = +0.0 and score + 0.0 * w == score
                    }
                }
                // ---- threshold compaction into the warp's staging buffer -------------------
                const bool keep = ok && score >= thr;
                const unsigned m = __ballot_sync(FULL_MASK, keep);
                if (m) {
                    const uint32_t n = __popc(m), slot = out_n + __popc(m & lanemask_lt());
                    st_kept += keep ? 1u : 0u;
                    if (coded) {
                        // unit-local 16-bit position + 16-bit score code; uncoded pairs as nsm_pair_t
                        const uint32_t code = keep ? dict_code(score) : 0u;
                        const bool exc = keep && code == 0xffffffffu, in = keep && !exc;
                        emit_pairs(exc, c_l, c_r, score, p.job.out_exc, p.job.out_exc_capacity,
                                   reinterpret_cast<unsigned long long *>(p.job.out_exc_count), p.job.out_flags);
                        const unsigned mi = __ballot_sync(FULL_MASK, in);
                        const uint32_t ni = __popc(mi), sl = out_n + __popc(mi & lanemask_lt());
                        nsm_cpacket_t &pk = s.out[warp].cpacket;
                        const uint32_t rec = (code << 16) | (li << 7) | rc;
                        if (in && sl < (uint32_t)NSM_CPACKET_RECORDS) pk.rec[sl] = rec;
                        if (out_n + ni >= (uint32_t)NSM_CPACKET_RECORDS) {
                            __syncwarp();
                            flush_cpacket(NSM_CPACKET_RECORDS, l0, r0);
                            if (in && sl >= (uint32_t)NSM_CPACKET_RECORDS) pk.rec[sl - NSM_CPACKET_RECORDS] = rec;
                            out_n = out_n + ni - NSM_CPACKET_RECORDS;
                        } else {
                            out_n += ni;
                        }
                    } else if (packets) {
                        // unit-local 16-bit position + score; the packet is filled to its last slot
                        nsm_packet_t &pk = s.out[warp].packet;
                        const uint16_t loc = (uint16_t)((li << 7) | rc);
                        if (keep && slot < (uint32_t)NSM_PACKET_RECORDS) { pk.score[slot] = score; pk.local[slot] = loc; }
                        if (out_n + n >= (uint32_t)NSM_PACKET_RECORDS) {
                            __syncwarp();
                            flush_packet(NSM_PACKET_RECORDS, l0, r0);
                            if (keep && slot >= (uint32_t)NSM_PACKET_RECORDS) {
                                pk.score[slot - NSM_PACKET_RECORDS] = score;
                                pk.local[slot - NSM_PACKET_RECORDS] = loc;
                            }
                            out_n = out_n + n - NSM_PACKET_RECORDS;
                        } else {
                            out_n += n;
                        }
                    } else {
                        if (out_n + n > J_OUT) flush_out();
                        if (keep) {
                            double2 rec;
                            rec.x = __longlong_as_double((long long)(((unsigned long long)c_r << 32) | c_l));
                            rec.y = score;
                            reinterpret_cast<double2 *>(&s.out[warp].pairs[0])[out_n + __popc(m & lanemask_lt())] = rec;
                        }
                        out_n += n;
                    }
                }
                __syncwarp();
            }
        }
        if ((packets || coded) && out_n) {  // positions are unit-local: the warp's partial packet ends here
            __syncwarp();
            if (coded) flush_cpacket(out_n, l0, r0); else flush_packet(out_n, l0, r0);
            out_n = 0;
        }
        unit = unit_after;
    }
    if (!packets && !coded) flush_out();

    if (p.job.out_stats) {
        atomicAdd(&s.stats[NSM_STAT_CANDIDATES], st_cand);
        atomicAdd(&s.stats[NSM_STAT_LEVEL_EVALS], st_evals);
        atomicAdd(&s.stats[NSM_STAT_LEVEL_MERGES], st_merges);
        atomicAdd(&s.stats[NSM_STAT_BOUND_PAIRS], st_bound);
        atomicAdd(&s.stats[NSM_STAT_KEPT], st_kept);
        __syncthreads();
        if (tid < NSM_N_STATS && s.stats[tid])
            atomicAdd(reinterpret_cast<unsigned long long *>(p.job.out_stats) + tid, s.stats[tid]);
    }
}

template <bool DEEP, int SPLIT, bool TWO = false>
static int launch_jaccard(const JaccardParams &p_in, uint64_t n_units, cudaStream_t stream) {
    JaccardParams p = p_in;
    p.unit_counter = NSM_J_DYN_UNITS ? next_unit_counter(stream) : nullptr;
    if (NSM_J_DYN_UNITS && !p.unit_counter) {
        set_error("unit counter: %s", cudaGetErrorString(cudaGetLastError()));
        return NSM_ERR_CUDA;
    }
    const size_t smem = sizeof(JaccardSmem);
    NSM_CUDA_CHECK(cudaFuncSetAttribute(jaccard_allpairs_kernel<DEEP, SPLIT, TWO>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t resident = (uint64_t)JT_CTAS * (uint64_t)sm_count();
    const uint32_t grid = (uint32_t)(n_units < resident ? n_units : resident);
    jaccard_allpairs_kernel<DEEP, SPLIT, TWO><<<grid, JT_THREADS, smem, stream>>>(p);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

template <bool DEEP>
static int launch_jaccard_split(const JaccardParams &p, uint64_t n_units, cudaStream_t stream) {
    switch (p.bound_split) {
        case 1: return p.two_small != NO_TWO ? launch_jaccard<DEEP, 1, true>(p, n_units, stream)
                                             : launch_jaccard<DEEP, 1>(p, n_units, stream);
        case 2: return launch_jaccard<DEEP, 2>(p, n_units, stream);
        case 0: if (!DEEP) return launch_jaccard<false, 0>(p, n_units, stream);  // coarse first half
                return launch_jaccard<DEEP, J_UNROLL>(p, n_units, stream);
        default: return launch_jaccard<DEEP, J_UNROLL>(p, n_units, stream);
    }
}

}  // namespace nsm

extern "C" int nsm_jaccard_allpairs(const nsm_sets_t *left, const nsm_sets_t *right,
                                    const nsm_job_t *job, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!left || !right || !job) { set_error("null argument"); return NSM_ERR_BAD_ARG; }
    if (int rc = prepare_job(job, left->n_items, stream)) return rc;
    if (job->out_mode != NSM_OUT_PAIRS && !job->out_stats) {
        set_error("the packet modes need out_stats (the kept-pair count is NSM_STAT_KEPT)");
        return NSM_ERR_BAD_ARG;
    }
    if (job->l_row_begin == job->l_row_end || right->n_items == 0) return NSM_OK;
    if (job->flat && (left->max_levels > 1 || right->max_levels > 1)) {
        set_error("flat scoring needs items with exactly one level");
        return NSM_ERR_BAD_ARG;
    }
    if (left->max_levels > 0xffffu || right->max_levels > 0xffffu) {
        set_error("items with more than 65535 levels are not supported");
        return NSM_ERR_UNSUPPORTED;
    }
    if ((uint64_t)left->n_slots * left->slot_stride > 0xffffffffull ||
        (uint64_t)right->n_slots * right->slot_stride > 0xffffffffull) {
        set_error("slot arrays beyond 2^32 entries are not supported");
        return NSM_ERR_UNSUPPORTED;
    }
    if (left->n_slots < 1 || left->n_slots > (uint32_t)J_SLOTS || right->n_slots < 1 ||
        right->n_slots > (uint32_t)J_SLOTS) {
        set_error("n_slots must be in 1..%d", J_SLOTS);
        return NSM_ERR_BAD_ARG;
    }

    JaccardParams p;
    p.L = *left; p.R = *right; p.job = *job;
    p.thr_lo = filter_threshold(job->threshold);
    // an item whose levels do not all fit its side's slots needs the CSR arrays for deep steps
    const bool l_deep = left->max_levels > left->n_slots + 1;
    const bool r_deep = right->max_levels > right->n_slots + 1;
    // Depth D: the smallest D such that steps beyond D cannot lift a zero score to the threshold,
    // 2^-D - 2^-Kmax < threshold with Kmax the deepest schedule of any pair.  Stage A uses it if
    // both sides hold D steps in their slots (or all their levels), otherwise the packed all-level
    // union; stage B tests its bound for the first time after D steps.
    p.any_depth = 0;
    p.bound_split = J_UNROLL;
    p.two_small = NO_TWO;
    p.dyn_tiles = 0;
    if (p.thr_lo > 0.0f && !job->flat) {
        const uint32_t kmax = left->max_levels > right->max_levels ? left->max_levels : right->max_levels;
        const float w_last = kmax < 120 ? ldexpf(1.0f, -(int)kmax) : 0.0f;
        uint32_t d = 1;
        while (d < 64 && !(ldexpf(1.0f, -(int)d) - w_last < p.thr_lo)) ++d;  // exact: powers of two
        if ((d <= left->n_slots || !l_deep) && (d <= right->n_slots || !r_deep))
            p.any_depth = d < (uint32_t)J_SLOTS ? d : (uint32_t)J_SLOTS;
        // splitting pays when few pairs survive the first half, i.e. at high thresholds (small D)
        p.bound_split = d <= 2 ? d : (uint32_t)J_UNROLL;
        // low thresholds (D >= 3) over nested levels that all lie in the slots: a coarse first half
        // (bound_split 0) proves most pairs that share nothing through step 2 below the threshold
        // deep funnels (D >= 3): stages B and C dominate and their load differs from warp to warp
        p.dyn_tiles = (NSM_J_DYN && d >= 3) ? 1u : 0u;
        if (NSM_J_COARSE && d >= 3 && left->nested && right->nested && !l_deep && !r_deep)
            p.bound_split = 0;
        // D == 1: score <= J_1 / 2 + (1/2 - 2^-Kmax), so a kept pair has J_1 >= jmin.  With at most
        // ONE shared id J_1 <= 1 / (a + b - 1), which is below jmin once a + b > 1 + 1 / jmin: if
        // both step-1 levels hold more than c = floor(floor(1 + 1/jmin) / 2) ids, a pair needs two
        // shared ids, hence (signature bits being injective for non-wild items) two shared bits.
        if (p.any_depth == 1 && p.bound_split == 1) {
            const double jmin = 2.0 * ((double)p.thr_lo - 0.5 + (double)w_last);
            if (jmin > 0.0) {
                const double t = floor((1.0 + 1.0 / jmin) * (1.0 + 1e-9));
                if (t < 64.0) p.two_small = (uint32_t)t / 2u;
            }
        }
    }
    const uint32_t n_rows = job->l_row_end - job->l_row_begin;
    p.n_lchunks = (n_rows + J_UNIT_LEFT - 1) / J_UNIT_LEFT;
    p.n_rblocks = (right->n_items + JT_THREADS - 1) / JT_THREADS;
    const uint64_t n_units = (uint64_t)p.n_lchunks * p.n_rblocks;
    if (n_units > 0xffffffffull) {
        set_error("too many work units (%llu); split the left row block", (unsigned long long)n_units);
        return NSM_ERR_UNSUPPORTED;
    }
    return (l_deep || r_deep) ? launch_jaccard_split<true>(p, n_units, stream)
                              : launch_jaccard_split<false>(p, n_units, stream);
}
