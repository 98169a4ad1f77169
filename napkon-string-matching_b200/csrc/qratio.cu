// fuzzy_match all-pairs kernel (sm_100a).
//
// Replaces the pair loop of ComparableData.gen_comparable for score_func == "fuzzy_match"
// (/root/reference/napkon_string_matching/types/comparable_data.py:223-243, compare_terms
// :248-265, compare/score_functions.py:20-27).  QRatio/100 of rapidfuzz 2.1.x is the normalised
// Indel similarity, 1 - (m + n - 2 LCS)/(m + n), so the kernel computes LCS
// lengths with Hyyro's bit-parallel recurrence  u = S & M[c];  S = (S + u) | (S - u)  on 64-bit
// words; LCS = number of zero bits of S.
//
// One pair per lane: every thread owns one right item and keeps the pattern-match masks M[c] of
// the right level string it is currently scoring in shared memory, transposed as [c][word][thread]
// so that the 32 lanes of a warp read 32 consecutive 64-bit words (no bank conflicts).  A unit of
// work is one block of right items x a group of left tiles; the left tile's level strings are
// staged in shared memory (8-byte aligned, read 8 characters at a time) and used as the text: all
// threads of the CTA consume the same character at the same time, so the inner loop has a uniform
// trip count, no divergence, and the mask loads of the next characters do not depend on S.
// Strings longer than 64 characters use W words per thread with the add carry chained through
// registers (W is a template parameter, <= 8, i.e. 512 characters).
//
// Two instantiations per W:
//   LEVELS = false  items have one level (the flat score function, the `Question` column of
//                   config 3): masks are built once per unit, every pair is scored and emitted
//                   as soon as its LCS is known.
//   LEVELS = true   compare_terms' schedule runs as the outer loop per left tile (step t uses
//                   level min(t, K-1) on both sides, weight 2^-t); a thread rebuilds its masks at
//                   most K_right times per tile; partial scores of the tile's pairs live in
//                   shared memory as float64 and are accumulated in the reference's order.
// Pairs with score >= threshold are compacted with one atomic per warp.
#include "nsm_common.cuh"

namespace nsm {

constexpr int Q_MAX_THREADS = 512;
constexpr int Q_TILE_FLAT = 32;       // left items per tile, one level each
constexpr int Q_TILE_LEVELS = 8;      // left items per tile when partial scores are kept
constexpr int Q_GROUP = 16;           // left tiles per unit
constexpr int Q_CHR_CAP = 16 * 1024;  // bytes of left level strings staged per tile
constexpr int Q_LEV_CAP = 1024;       // left levels staged per tile
constexpr int Q_MAX_WORDS = 8;
constexpr size_t Q_SMEM_BUDGET = 220 * 1024;

struct QratioParams {
    nsm_strings_t L, R;
    nsm_job_t job;
    uint32_t tile_left, n_ltiles, n_lgroups, n_rblocks;
    uint32_t threads;   // right items per block
    uint32_t n_alpha;   // rows of the mask table
    uint32_t r_begin, r_end;  // the right items of this launch (one word-count class)
};

struct QratioLayout {  // offsets into dynamic shared memory
    size_t pm, acc, chr, lev_off, lev_len, item_g0, cat, misc, total;
};

__host__ __device__ inline QratioLayout qratio_layout(uint32_t n_alpha, uint32_t words,
                                                      uint32_t threads, uint32_t tile_left,
                                                      bool levels) {
    QratioLayout l;
    size_t o = 0;
    l.pm = o;      o += (size_t)n_alpha * words * threads * 8;
    l.acc = o;     o += levels ? (size_t)tile_left * threads * 8 : 0;
    l.chr = o;     o += Q_CHR_CAP;
    l.cat = o;     o += Q_TILE_FLAT * 8;
    l.lev_off = o; o += Q_LEV_CAP * 4;
    l.lev_len = o; o += Q_LEV_CAP * 4;
    l.item_g0 = o; o += (Q_TILE_FLAT + 1) * 4;
    l.misc = o;    o += 16;
    l.total = (o + 15) & ~(size_t)15;
    return l;
}

// sum = S + u over W 64-bit words as ONE multi-word addition: the carry runs through the
// hardware carry flag (add.cc / addc.cc on the 32-bit halves), two instructions per word instead
// of an add plus compares and selects per word.  HALVES = 2 * words of one block (<= 8).
__device__ __forceinline__ uint32_t lo32(uint64_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi32(uint64_t v) { return (uint32_t)(v >> 32); }
__device__ __forceinline__ uint64_t mk64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// one block of two words; carry_in / carry_out are 0 or 1
__device__ __forceinline__ void add2(const uint64_t *a, const uint64_t *b, uint64_t *r, uint32_t cin,
                                     uint32_t &cout) {
    uint32_t r0, r1, r2, r3, t;  // t: scratch of the flag-setting add
    asm("{\n\t"
        "add.cc.u32 %5, %14, 0xffffffff;\n\t"   // carry flag = carry_in
        "addc.cc.u32 %0, %6, %10;\n\t"
        "addc.cc.u32 %1, %7, %11;\n\t"
        "addc.cc.u32 %2, %8, %12;\n\t"
        "addc.cc.u32 %3, %9, %13;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(cout), "=r"(t)
        : "r"(lo32(a[0])), "r"(hi32(a[0])), "r"(lo32(a[1])), "r"(hi32(a[1])),
          "r"(lo32(b[0])), "r"(hi32(b[0])), "r"(lo32(b[1])), "r"(hi32(b[1])), "r"(cin));
    (void)t;
    r[0] = mk64(r0, r1); r[1] = mk64(r2, r3);
}

// two words, no carry in, no carry out (the whole pattern of the W = 2 class)
__device__ __forceinline__ void add2_only(const uint64_t *a, const uint64_t *b, uint64_t *r) {
    uint32_t r0, r1, r2, r3;
    asm("{\n\t"
        "add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, %9;\n\t"
        "addc.cc.u32 %2, %6, %10;\n\t"
        "addc.u32 %3, %7, %11;\n\t"
        "}"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
        : "r"(lo32(a[0])), "r"(hi32(a[0])), "r"(lo32(a[1])), "r"(hi32(a[1])),
          "r"(lo32(b[0])), "r"(hi32(b[0])), "r"(lo32(b[1])), "r"(hi32(b[1])));
    r[0] = mk64(r0, r1); r[1] = mk64(r2, r3);
}

// one word with carry in and out
__device__ __forceinline__ void add1(uint64_t a, uint64_t b, uint64_t &r, uint32_t cin, uint32_t &cout) {
    uint32_t r0, r1, t;
    asm("{\n\t"
        "add.cc.u32 %3, %8, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %4, %6;\n\t"
        "addc.cc.u32 %1, %5, %7;\n\t"
        "addc.u32 %2, 0, 0;\n\t"
        "}"
        : "=r"(r0), "=r"(r1), "=r"(cout), "=r"(t)
        : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)), "r"(cin));
    r = mk64(r0, r1);
}

template <int W>
__device__ __forceinline__ void add_words(const uint64_t (&a)[W], const uint64_t (&b)[W], uint64_t (&r)[W]) {
    if (W == 1) {
        r[0] = a[0] + b[0];
    } else if (W == 2) {
        add2_only(a, b, r);
    } else {
        uint32_t carry = 0;
#pragma unroll
        for (int x = 0; x + 2 <= W; x += 2) add2(a + x, b + x, r + x, carry, carry);
        if (W & 1) add1(a[W - 1], b[W - 1], r[W - 1], carry, carry);
    }
}

// LCS length of my pattern (masks in shared memory, column `pm`) and a text of n characters that
// starts 8-byte aligned at `text` in shared memory.  Uniform over the CTA.
template <int W>
__device__ __forceinline__ uint32_t lcs_bitparallel(const uint64_t *__restrict__ pm, uint32_t nthr,
                                                    const uint8_t *__restrict__ text, uint32_t n) {
    uint64_t S[W];
#pragma unroll
    for (int x = 0; x < W; ++x) S[x] = ~0ull;
    const uint2 *text8 = reinterpret_cast<const uint2 *>(text);
    // byte addressing: the mask row of character c starts c * row_bytes after my column, so a
    // character costs one byte extract (PRMT), one multiply-add and the load
    const unsigned char *col = reinterpret_cast<const unsigned char *>(pm);
    const uint32_t word_bytes = nthr * 8u, row_bytes = (uint32_t)W * word_bytes;
    auto step = [&](uint32_t c) {
        const unsigned char *row = col + c * row_bytes;
        uint64_t M[W], u[W], sum[W];
#pragma unroll
        for (int x = 0; x < W; ++x) {
            M[x] = *reinterpret_cast<const uint64_t *>(row + (uint32_t)x * word_bytes);
            u[x] = S[x] & M[x];
        }
        add_words<W>(S, u, sum);
        // u is a subset of S, so S - u == S & ~M: one three-input logic op per half instead of a
        // subtract with borrow
#pragma unroll
        for (int x = 0; x < W; ++x) S[x] = sum[x] | (S[x] & ~M[x]);
    };
    uint32_t j = 0;
    for (; j + 8 <= n; j += 8) {  // eight characters per shared-memory word
        const uint2 w8 = text8[j >> 3];
#pragma unroll
        for (int q = 0; q < 4; ++q) step(__byte_perm(w8.x, 0u, 0x4440u + q));
#pragma unroll
        for (int q = 0; q < 4; ++q) step(__byte_perm(w8.y, 0u, 0x4440u + q));
    }
    if (j < n) {
        const uint2 w8 = text8[j >> 3];
        for (uint32_t q = 0; j + q < n; ++q)
            step(__byte_perm(q < 4 ? w8.x : w8.y, 0u, 0x4440u + (q & 3u)));
    }
    uint32_t lcs = 0;
#pragma unroll
    for (int x = 0; x < W; ++x) lcs += __popcll(~S[x]);
    return lcs;
}

// Two texts against my pattern at once: the two S chains are independent, so their dependent
// AND -> ADD -> OR sequences interleave and hide each other's latency (a CTA holds few warps when
// the mask tables are large).  The texts run in lockstep over the length of the shorter one; the
// rest of each is finished by itself.
template <int W>
__device__ __forceinline__ void lcs_bitparallel2(const uint64_t *__restrict__ pm, uint32_t nthr,
                                                 const uint8_t *__restrict__ text_a, uint32_t na,
                                                 const uint8_t *__restrict__ text_b, uint32_t nb,
                                                 uint32_t &lcs_a, uint32_t &lcs_b) {
    uint64_t Sa[W], Sb[W];
#pragma unroll
    for (int x = 0; x < W; ++x) Sa[x] = Sb[x] = ~0ull;
    const uint2 *a8 = reinterpret_cast<const uint2 *>(text_a), *b8 = reinterpret_cast<const uint2 *>(text_b);
    const unsigned char *col = reinterpret_cast<const unsigned char *>(pm);
    const uint32_t word_bytes = nthr * 8u, row_bytes = (uint32_t)W * word_bytes;
    auto step = [&](uint64_t (&S)[W], uint32_t c) {
        const unsigned char *row = col + c * row_bytes;
        uint64_t M[W], u[W], sum[W];
#pragma unroll
        for (int x = 0; x < W; ++x) {
            M[x] = *reinterpret_cast<const uint64_t *>(row + (uint32_t)x * word_bytes);
            u[x] = S[x] & M[x];
        }
        add_words<W>(S, u, sum);
#pragma unroll
        for (int x = 0; x < W; ++x) S[x] = sum[x] | (S[x] & ~M[x]);
    };
    const uint32_t nc = min(na, nb);
    uint32_t j = 0;
    for (; j + 8 <= nc; j += 8) {
        const uint2 wa = a8[j >> 3], wb = b8[j >> 3];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            step(Sa, __byte_perm(wa.x, 0u, 0x4440u + q));
            step(Sb, __byte_perm(wb.x, 0u, 0x4440u + q));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            step(Sa, __byte_perm(wa.y, 0u, 0x4440u + q));
            step(Sb, __byte_perm(wb.y, 0u, 0x4440u + q));
        }
    }
    auto finish = [&](uint64_t (&S)[W], const uint2 *t8, uint32_t n) {
        uint32_t i = j;
        for (; i + 8 <= n; i += 8) {
            const uint2 w8 = t8[i >> 3];
#pragma unroll
            for (int q = 0; q < 4; ++q) step(S, __byte_perm(w8.x, 0u, 0x4440u + q));
#pragma unroll
            for (int q = 0; q < 4; ++q) step(S, __byte_perm(w8.y, 0u, 0x4440u + q));
        }
        if (i < n) {
            const uint2 w8 = t8[i >> 3];
            for (uint32_t q = 0; i + q < n; ++q)
                step(S, __byte_perm(q < 4 ? w8.x : w8.y, 0u, 0x4440u + (q & 3u)));
        }
    };
    finish(Sa, a8, na);
    finish(Sb, b8, nb);
    lcs_a = lcs_b = 0;
#pragma unroll
    for (int x = 0; x < W; ++x) { lcs_a += __popcll(~Sa[x]); lcs_b += __popcll(~Sb[x]); }
}

#ifndef Q_DUAL
#define Q_DUAL 1   // score two left strings per round where the registers allow it (W <= 4)
#endif

// Reciprocals of the possible length sums (two strings of up to 64 * Q_MAX_WORDS characters).
struct LenRcpTable {
    double v[64 * Q_MAX_WORDS * 2 + 1];
    constexpr LenRcpTable() : v() {
        for (int u = 1; u <= 64 * Q_MAX_WORDS * 2; ++u) v[u] = 1.0 / (double)u;
    }
};
__device__ const LenRcpTable g_len_rcp = LenRcpTable();

// a / b as  q0 = RN(a * r);  q = fma(fma(-q0, b, a), r, q0)  with r = RN(1 / b).  For the operands
// qratio_from_lcs feeds it this equals the correctly rounded quotient - every case is enumerated
// by tests/csrc/fast_ratio_check.c - and it costs three float64 operations instead of the general
// division sequence.
__device__ __forceinline__ double div_by_rcp(double a, double b, double r) {
    const double q0 = __dmul_rn(a, r);
    return __fma_rn(__fma_rn(-q0, b, a), r, q0);
}

// QRatio(a, b) / 100 from the counts: 0 when either processed string is empty, else
// ((1.0 - dist / lensum) * 100) / 100 with dist = lensum - 2 LCS — this exact operation order.
__device__ __forceinline__ double qratio_from_lcs(uint32_t m, uint32_t n, uint32_t lcs) {
    if (m == 0 || n == 0) return 0.0;
    const uint32_t lensum = m + n, dist = lensum - 2u * lcs;
    if (lensum <= 64u * Q_MAX_WORDS * 2u) {
        const double norm_dist = div_by_rcp((double)dist, (double)lensum, g_len_rcp.v[lensum]);
        const double norm_sim = __dsub_rn(1.0, norm_dist);
        return div_by_rcp(__dmul_rn(norm_sim, 100.0), 100.0, 0.01);  // 0.01 is RN(1 / 100)
    }
    const double norm_dist = __ddiv_rn((double)dist, (double)lensum);
    const double norm_sim = __dsub_rn(1.0, norm_dist);
    return __ddiv_rn(__dmul_rn(norm_sim, 100.0), 100.0);
}

template <int W, bool LEVELS>
__global__ void __launch_bounds__(Q_MAX_THREADS, 1)
qratio_allpairs_kernel(const QratioParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x, nthr = blockDim.x;
    const QratioLayout lay = qratio_layout(p.n_alpha, W, nthr, p.tile_left, LEVELS);
    uint64_t *s_pm = reinterpret_cast<uint64_t *>(smem_raw + lay.pm);      // [c][w][thread]
    double *s_acc = reinterpret_cast<double *>(smem_raw + lay.acc);        // [li][thread]
    uint8_t *s_chr = smem_raw + lay.chr;
    uint32_t *s_lev_off = reinterpret_cast<uint32_t *>(smem_raw + lay.lev_off);
    uint32_t *s_lev_len = reinterpret_cast<uint32_t *>(smem_raw + lay.lev_len);
    uint32_t *s_item_g0 = reinterpret_cast<uint32_t *>(smem_raw + lay.item_g0);
    uint64_t *s_cat = reinterpret_cast<uint64_t *>(smem_raw + lay.cat);
    uint32_t *s_misc = reinterpret_cast<uint32_t *>(smem_raw + lay.misc);  // [0] max K right

    unsigned long long *count = reinterpret_cast<unsigned long long *>(p.job.out_count);
    const bool flat = p.job.flat != 0;
    const double thr = p.job.threshold;
    unsigned long long st_evals = 0;
    const uint64_t *pm = s_pm + tid;

    // (re)build my pattern masks for right level g; returns its length
    auto build_masks = [&](uint32_t g) {
        const uint32_t c0 = __ldg(p.R.level_chr_off + g), m = __ldg(p.R.level_len + g);
        for (uint32_t row = 0; row < p.n_alpha * W; ++row) s_pm[(size_t)row * nthr + tid] = 0;
        for (uint32_t j = 0; j < m; j += 8) {
            const uint64_t w8 = __ldg(reinterpret_cast<const uint64_t *>(p.R.chr + c0 + j));
            for (uint32_t q = 0; q < 8 && j + q < m; ++q) {
                const uint32_t c = (uint32_t)(w8 >> (8 * q)) & 0xffu;
                s_pm[((size_t)c * W + ((j + q) >> 6)) * nthr + tid] |= 1ull << ((j + q) & 63u);
            }
        }
        return m;
    };

    const uint32_t n_units = p.n_lgroups * p.n_rblocks;
    for (uint32_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const uint32_t rb = unit / p.n_lgroups, lgroup = unit - rb * p.n_lgroups;
        const uint32_t r = p.r_begin + rb * nthr + tid;
        const bool r_valid = r < p.r_end;
        uint32_t rg0 = 0, kr = 0;
        uint64_t rcat = 0;
        if (r_valid) {
            rg0 = __ldg(p.R.item_level_off + r);
            kr = __ldg(p.R.item_level_off + r + 1) - rg0;
            if (p.job.cat_mode) rcat = __ldg(p.job.r_cat + r);
        }
        uint32_t cur_slot = 0xffffffffu, m = 0;
        if (!LEVELS && kr) { m = build_masks(rg0); cur_slot = 0; }  // my columns only: no barrier needed
        if (LEVELS) {
            __syncthreads();  // s_misc of the previous unit consumed
            if (tid == 0) s_misc[0] = 0;
            __syncthreads();
            uint32_t k = kr;
            for (int o = 16; o; o >>= 1) k = max(k, __shfl_xor_sync(FULL_MASK, k, o));
            if ((tid & 31u) == 0) atomicMax(&s_misc[0], k);
        }

        const uint32_t lt_begin = lgroup * Q_GROUP;
        const uint32_t lt_end = min(lt_begin + (uint32_t)Q_GROUP, p.n_ltiles);
        for (uint32_t lt = lt_begin; lt < lt_end; ++lt) {
            const uint32_t l0 = p.job.l_row_begin + lt * p.tile_left;
            const uint32_t nl = min(p.tile_left, p.job.l_row_end - l0);
            const uint32_t G0 = __ldg(p.L.item_level_off + l0);
            const uint32_t nlev = __ldg(p.L.item_level_off + l0 + nl) - G0;
            uint32_t C0 = 0, nchr = 0;
            if (nlev) {
                C0 = __ldg(p.L.level_chr_off + G0);
                nchr = __ldg(p.L.level_chr_off + G0 + nlev - 1) +
                       ((__ldg(p.L.level_len + G0 + nlev - 1) + 7u) & ~7u) - C0;
            }

            __syncthreads();  // previous tile fully consumed
            for (uint32_t i = tid; i < nchr / 8; i += nthr)
                reinterpret_cast<uint64_t *>(s_chr)[i] =
                    __ldg(reinterpret_cast<const uint64_t *>(p.L.chr + C0) + i);
            for (uint32_t g = tid; g < nlev; g += nthr) {
                s_lev_off[g] = __ldg(p.L.level_chr_off + G0 + g) - C0;
                s_lev_len[g] = __ldg(p.L.level_len + G0 + g);
            }
            // strided: a launch may run fewer threads than the tile has items (small right sides)
            for (uint32_t i = tid; i <= nl; i += nthr) s_item_g0[i] = __ldg(p.L.item_level_off + l0 + i) - G0;
            for (uint32_t i = tid; i < nl; i += nthr) s_cat[i] = p.job.cat_mode ? __ldg(p.job.l_cat + l0 + i) : 0;
            __syncthreads();

            if (!LEVELS) {
                // ---- one level per item: score and emit pair by pair ------------------------
                auto emit_flat = [&](uint32_t li, uint32_t kl, uint32_t n, uint32_t lcs) {
                    bool ok = r_valid && keep_categories(p.job.cat_mode, s_cat[li], rcat);
                    double score = 0.0;
                    if (kl && kr) {
                        // flat: score_func(l0, r0); else compare_terms on K = 1 items: t = 1, weight 1/2
                        score = qratio_from_lcs(m, n, lcs);
                        if (!flat) score = __fma_rn(score, 0.5, 0.0);
                        ++st_evals;
                    }
                    if (ok && (kl == 0) != (kr == 0)) {  // IndexError in the reference
                        atomicOr(p.job.out_flags, NSM_FLAG_EMPTY_ITEM);
                        ok = false;
                    }
                    emit_pairs(ok && score >= thr, l0 + li, r, score, static_cast<nsm_pair_t *>(p.job.out_pairs),
                               p.job.out_capacity, count, p.job.out_flags);
                };
                // (uniform control flow; threads without a right level compute on stale masks, unused)
                for (uint32_t li = 0; li < nl;) {
                    const uint32_t lg0 = s_item_g0[li], kl = s_item_g0[li + 1] - lg0;
                    const uint32_t n = kl ? s_lev_len[lg0] : 0u;
                    if (Q_DUAL && W <= 4 && li + 1 < nl) {
                        const uint32_t lg1 = s_item_g0[li + 1], kl1 = s_item_g0[li + 2] - lg1;
                        if (kl && kl1) {
                            const uint32_t n1 = s_lev_len[lg1];
                            uint32_t lcs0, lcs1;
                            lcs_bitparallel2<W>(pm, nthr, s_chr + s_lev_off[lg0], n, s_chr + s_lev_off[lg1], n1,
                                                lcs0, lcs1);
                            emit_flat(li, kl, n, lcs0);
                            emit_flat(li + 1, kl1, n1, lcs1);
                            li += 2;
                            continue;
                        }
                    }
                    const uint32_t lcs = kl ? lcs_bitparallel<W>(pm, nthr, s_chr + s_lev_off[lg0], n) : 0u;
                    emit_flat(li, kl, n, lcs);
                    ++li;
                }
            } else {
                // ---- compare_terms' schedule as the outer loop, partial scores in smem --------
                for (uint32_t li = 0; li < nl; ++li) s_acc[(size_t)li * nthr + tid] = 0.0;
                uint32_t max_kl = 0;
                for (uint32_t li = 0; li < nl; ++li) max_kl = max(max_kl, s_item_g0[li + 1] - s_item_g0[li]);
                const uint32_t max_kr = s_misc[0];
                const uint32_t t_end = max(max_kr, max_kl);
                double w = 1.0;
                for (uint32_t t = 1; t <= t_end; ++t) {
                    w *= 0.5;
                    if (kr) {
                        const uint32_t slot = min(t, kr - 1);
                        if (slot != cur_slot) { cur_slot = slot; m = build_masks(rg0 + slot); }
                    }
                    // the tile's items that take part in step t (uniform), two per round
                    auto level_of = [&](uint32_t li, uint32_t &kl) {
                        const uint32_t lg0 = s_item_g0[li];
                        kl = s_item_g0[li + 1] - lg0;
                        return lg0 + min(t, max(kl, 1u) - 1);
                    };
                    auto takes_part = [&](uint32_t li) {
                        const uint32_t kl = s_item_g0[li + 1] - s_item_g0[li];
                        return kl != 0 && t <= max(kl, max_kr);
                    };
                    auto add_score = [&](uint32_t li, uint32_t kl, uint32_t n, uint32_t lcs) {
                        if (r_valid && kr && t <= max(kl, kr)) {
                            double *a = s_acc + (size_t)li * nthr + tid;
                            *a = __fma_rn(qratio_from_lcs(m, n, lcs), w, *a);
                            ++st_evals;
                        }
                    };
                    uint32_t li = 0;
                    while (true) {
                        while (li < nl && !takes_part(li)) ++li;
                        if (li >= nl) break;
                        uint32_t kl0, kl1 = 0;
                        const uint32_t g0 = level_of(li, kl0), n0 = s_lev_len[g0];
                        uint32_t lj = li + 1;
                        if (Q_DUAL && W <= 4) {
                            while (lj < nl && !takes_part(lj)) ++lj;
                        } else {
                            lj = nl;
                        }
                        if (lj < nl) {
                            const uint32_t g1 = level_of(lj, kl1), n1 = s_lev_len[g1];
                            uint32_t lcs0, lcs1;
                            lcs_bitparallel2<W>(pm, nthr, s_chr + s_lev_off[g0], n0, s_chr + s_lev_off[g1], n1,
                                                lcs0, lcs1);
                            add_score(li, kl0, n0, lcs0);
                            add_score(lj, kl1, n1, lcs1);
                            li = lj + 1;
                        } else {
                            const uint32_t lcs = lcs_bitparallel<W>(pm, nthr, s_chr + s_lev_off[g0], n0);
                            add_score(li, kl0, n0, lcs);
                            li = (Q_DUAL && W <= 4) ? nl : li + 1;
                        }
                    }
                }
                cur_slot = 0xffffffffu;  // the next tile starts over at step 1
                for (uint32_t li = 0; li < nl; ++li) {
                    const uint32_t kl = s_item_g0[li + 1] - s_item_g0[li];
                    bool ok = r_valid && keep_categories(p.job.cat_mode, s_cat[li], rcat);
                    if (ok && (kl == 0) != (kr == 0)) {  // IndexError in the reference
                        atomicOr(p.job.out_flags, NSM_FLAG_EMPTY_ITEM);
                        ok = false;
                    }
                    const double score = s_acc[(size_t)li * nthr + tid];
                    emit_pairs(ok && score >= thr, l0 + li, r, score, static_cast<nsm_pair_t *>(p.job.out_pairs),
                               p.job.out_capacity, count, p.job.out_flags);
                }
            }
        }
    }
    if (p.job.out_stats) {
        for (int o = 16; o; o >>= 1) st_evals += __shfl_xor_sync(FULL_MASK, st_evals, o);
        if ((tid & 31u) == 0 && st_evals)
            atomicAdd(reinterpret_cast<unsigned long long *>(p.job.out_stats) + NSM_STAT_LEVEL_EVALS,
                      st_evals);
    }
}

template <int W, bool LEVELS>
static int launch_qratio(const QratioParams &p, size_t smem, uint32_t grid, cudaStream_t stream) {
    NSM_CUDA_CHECK(cudaFuncSetAttribute(qratio_allpairs_kernel<W, LEVELS>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qratio_allpairs_kernel<W, LEVELS><<<grid, p.threads, smem, stream>>>(p);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

template <bool LEVELS>
static int dispatch_qratio(uint32_t words, const QratioParams &p, size_t smem, uint32_t grid,
                           cudaStream_t stream) {
    switch (words) {
        case 1: return launch_qratio<1, LEVELS>(p, smem, grid, stream);
        case 2: return launch_qratio<2, LEVELS>(p, smem, grid, stream);
        case 3: return launch_qratio<3, LEVELS>(p, smem, grid, stream);
        case 4: return launch_qratio<4, LEVELS>(p, smem, grid, stream);
        case 6: return launch_qratio<6, LEVELS>(p, smem, grid, stream);
        default: return launch_qratio<8, LEVELS>(p, smem, grid, stream);
    }
}

}  // namespace nsm

extern "C" int nsm_qratio_allpairs(const nsm_strings_t *left, const nsm_strings_t *right,
                                   const nsm_job_t *job, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!left || !right || !job) { set_error("null argument"); return NSM_ERR_BAD_ARG; }
    if (int rc = prepare_job(job, left->n_items, stream)) return rc;
    if (job->out_mode != NSM_OUT_PAIRS) {
        set_error("nsm_qratio_allpairs writes nsm_pair_t records only (out_mode NSM_OUT_PAIRS)");
        return NSM_ERR_UNSUPPORTED;
    }
    if (job->l_row_begin == job->l_row_end || right->n_items == 0) return NSM_OK;
    if (job->flat && (left->max_levels > 1 || right->max_levels > 1)) {
        set_error("flat scoring needs items with exactly one level");
        return NSM_ERR_BAD_ARG;
    }
    if (left->n_alphabet != right->n_alphabet || left->n_alphabet > 255) {
        set_error("both sides must be packed with one alphabet of <= 255 codes");
        return NSM_ERR_BAD_ARG;
    }
    if (right->class_end[Q_MAX_WORDS - 1] != right->n_items) {
        set_error("right level strings of up to %u characters; the kernel handles %d",
                  right->max_len, 64 * Q_MAX_WORDS);
        return NSM_ERR_UNSUPPORTED;
    }
    const bool levels = left->max_levels > 1 || right->max_levels > 1;
    const uint32_t kl = left->max_levels ? left->max_levels : 1u;
    const uint32_t per_item_chr = kl * (((left->max_len ? left->max_len : 1u) + 7u) & ~7u);
    uint32_t tl = levels ? (uint32_t)Q_TILE_LEVELS : (uint32_t)Q_TILE_FLAT;
    if (per_item_chr * tl > (uint32_t)Q_CHR_CAP) tl = (uint32_t)Q_CHR_CAP / per_item_chr;
    if (kl * tl > (uint32_t)Q_LEV_CAP) tl = (uint32_t)Q_LEV_CAP / kl;
    if (tl == 0) {
        set_error("a left item (%u levels x %u characters) exceeds the staged tile", kl, left->max_len);
        return NSM_ERR_UNSUPPORTED;
    }

    QratioParams p;
    p.L = *left; p.R = *right; p.job = *job;
    p.tile_left = tl;
    p.n_alpha = left->n_alphabet ? left->n_alphabet : 1u;
    const uint32_t n_rows = job->l_row_end - job->l_row_begin;
    p.n_ltiles = (n_rows + tl - 1) / tl;
    p.n_lgroups = (p.n_ltiles + Q_GROUP - 1) / Q_GROUP;

    // one launch per word-count class of the right side (its items are stored class by class),
    // each with the narrowest bit-vectors and as many threads as its mask tables leave room for
    for (uint32_t w = 0; w < (uint32_t)Q_MAX_WORDS; ++w) {
        p.r_begin = w ? right->class_end[w - 1] : 0u;
        p.r_end = right->class_end[w];
        if (p.r_end <= p.r_begin) continue;
        const uint32_t words = w + 1;
        // words actually instantiated: 1, 2, 3, 4, 6, 8
        const uint32_t w_inst = words <= 4 ? words : (words <= 6 ? 6u : 8u);
        uint32_t threads = Q_MAX_THREADS;
        while (threads >= 32 && qratio_layout(p.n_alpha, w_inst, threads, tl, levels).total > Q_SMEM_BUDGET)
            threads -= 32;
        if (threads < 32) {
            set_error("alphabet %u x %u words does not fit shared memory", p.n_alpha, w_inst);
            return NSM_ERR_UNSUPPORTED;
        }
        // no more threads than right items (rounded up to a warp): a small right side, e.g. one
        // term against many synonyms, should not pay for idle mask columns
        const uint32_t n_right = p.r_end - p.r_begin;
        const uint32_t need = ((n_right + 31u) / 32u) * 32u;
        if (threads > need) threads = need;
        p.threads = threads;
        const size_t smem = qratio_layout(p.n_alpha, w_inst, threads, tl, levels).total;
        p.n_rblocks = (n_right + threads - 1) / threads;
        const uint64_t n_units = (uint64_t)p.n_lgroups * p.n_rblocks;
        if (n_units > 0xffffffffull) {
            set_error("too many work units (%llu); split the left row block", (unsigned long long)n_units);
            return NSM_ERR_UNSUPPORTED;
        }
        const uint32_t resident = (uint32_t)sm_count();
        const uint32_t grid = (uint32_t)(n_units < resident ? n_units : resident);
        const int rc = levels ? dispatch_qratio<true>(w_inst, p, smem, grid, stream)
                              : dispatch_qratio<false>(w_inst, p, smem, grid, stream);
        if (rc) return rc;
    }
    return NSM_OK;
}
