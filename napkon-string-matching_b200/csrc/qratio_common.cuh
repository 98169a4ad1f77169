// Shared pieces of the fuzzy_match kernels (sm_100a): Hyyro's bit-parallel LCS recurrence on
// 64-bit words with the add carried through the hardware carry chain, and the exact
// int -> float64 map of QRatio / 100 (see qratio.cu for the references).
#pragma once

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
    uint32_t swap_out;        // 1: emit (right, left): the caller swapped the sides
    uint32_t *unit_counter;   // device counter the CTAs draw their units from; NULL: fixed stride
};

// qratio_flat.cu: items with one level each; right items [r_lo, r_hi) of the classes <= 8 words
int qratio_flat_launch(const nsm_strings_t *left, const nsm_strings_t *right, const nsm_job_t *job,
                       uint32_t r_lo, uint32_t r_hi, bool swap_out, cudaStream_t stream);
// qratio_long.cu: any lengths, one warp per item pair
int qratio_long_launch(const nsm_strings_t *left, const nsm_strings_t *right, const nsm_job_t *job,
                       uint32_t l_begin, uint32_t l_end, uint32_t r_begin, uint32_t r_end, bool swap_out,
                       cudaStream_t stream);

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
static __device__ const LenRcpTable g_len_rcp = LenRcpTable();

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


}  // namespace nsm
