// fuzzy_match for level strings of ANY length (sm_100a): one warp per item pair, the bit-vector of
// Hyyro's LCS recurrence spread over the lanes ("warp shuffles for strings longer than 64
// characters", north_star) — lane l holds the 64-bit words l, l + 32, ... of the pattern, and the
// carry of the multi-word addition S + (S & M) crosses the lanes through two ballots: with G =
// lanes whose word overflowed and P = lanes whose word became all ones, the carries INTO the lanes
// are the carry bits of the integer addition (G | P) + G (+ carry in), i.e. sum ^ (G | P) ^ G.
//
// The kernels of qratio.cu / qratio_flat.cu keep one pattern per THREAD and stop at 8 words (512
// characters); pairs whose two items both hold a longer level string — rare: questionnaire texts
// are tens of characters — come here, so that no input the reference scores
// (compare/score_functions.py:20-27 has no length limit) is refused.  The shorter string of an
// evaluation is the pattern (LCS is symmetric); its pattern-match masks live in shared memory as
// [code][word].  compare_terms' schedule (comparable_data.py:248-265) runs per pair as in the
// reference: step t scores level min(t, K-1) of both items with weight 2^-t, accumulated in order.
#include "qratio_common.cuh"

namespace nsm {

constexpr int QL_WORDS_PER_LANE = 8;                       // 8 x 32 words = 16384 characters
constexpr uint32_t QL_MAX_WORDS = 32u * QL_WORDS_PER_LANE;
constexpr size_t QL_SMEM_BUDGET = 200 * 1024;

struct QlongParams {
    nsm_strings_t L, R;
    nsm_job_t job;
    uint32_t l_begin, l_end, r_begin, r_end;  // the item ranges of this launch
    uint32_t words_cap;                       // words per mask row the shared table is sized for
    uint32_t swap_out;
};

__global__ void __launch_bounds__(32, 1) qratio_long_kernel(const QlongParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *tab = reinterpret_cast<uint64_t *>(smem_raw);   // [code][words of this evaluation]
    const unsigned lane = threadIdx.x;
    unsigned long long *count = reinterpret_cast<unsigned long long *>(p.job.out_count);
    nsm_pair_t *out = static_cast<nsm_pair_t *>(p.job.out_pairs);
    const bool flat = p.job.flat != 0;
    const double thr = p.job.threshold;
    const uint32_t n_alpha = p.L.n_alphabet ? p.L.n_alphabet : 1u;
    const uint64_t n_l = p.l_end - p.l_begin, n_r = p.r_end - p.r_begin;
    unsigned long long st_evals = 0;

    // LCS length of two level strings, computed by the whole warp (arguments warp-uniform)
    auto lcs_warp = [&](const uint8_t *a, uint32_t na, const uint8_t *b, uint32_t nb) -> uint32_t {
        const uint8_t *pat = a, *txt = b;
        uint32_t mp = na, nt = nb;
        if (mp > nt) { pat = b; txt = a; mp = nb; nt = na; }
        const uint32_t W = (mp + 63u) >> 6;
        __syncwarp();
        for (uint32_t i = lane; i < n_alpha * W; i += 32) tab[i] = 0;
        __syncwarp();
        for (uint32_t w = lane; w < W; w += 32) {   // word w is built by one lane: no atomics
            const uint32_t end = min(mp, (w + 1u) << 6);
            for (uint32_t i = w << 6; i < end; ++i) tab[(uint32_t)__ldg(pat + i) * W + w] |= 1ull << (i & 63u);
        }
        __syncwarp();
        uint64_t S[QL_WORDS_PER_LANE];
#pragma unroll
        for (int k = 0; k < QL_WORDS_PER_LANE; ++k) S[k] = ~0ull;
        const uint32_t groups = (W + 31u) >> 5;
        for (uint32_t j = 0; j < nt; ++j) {
            const uint64_t *row = tab + (uint32_t)__ldg(txt + j) * W;
            uint64_t carry = 0;   // into lane 0 of the next group of 32 words
#pragma unroll
            for (int k = 0; k < QL_WORDS_PER_LANE; ++k) {
                if ((uint32_t)k < groups) {   // warp-uniform
                    const uint32_t w = ((uint32_t)k << 5) | lane;
                    const bool on = w < W;
                    const uint64_t M = on ? row[w] : 0ull;
                    const uint64_t u = S[k] & M, s = S[k] + u;
                    const uint64_t G = __ballot_sync(FULL_MASK, on && s < u);
                    const uint64_t P = __ballot_sync(FULL_MASK, on && s == ~0ull);
                    const uint64_t X = G | P, sum = X + G + carry;
                    const uint64_t cin = sum ^ X ^ G;           // bit i: carry into lane i
                    carry = (cin >> 32) & 1ull;
                    if (on) S[k] = (s + ((cin >> lane) & 1ull)) | (S[k] & ~M);
                }
            }
        }
        uint32_t lcs = 0;
#pragma unroll
        for (int k = 0; k < QL_WORDS_PER_LANE; ++k)
            if ((((uint32_t)k << 5) | lane) < W) lcs += __popcll(~S[k]);
        return __reduce_add_sync(FULL_MASK, lcs);
    };

    for (uint64_t pair = blockIdx.x; pair < n_l * n_r; pair += gridDim.x) {
        const uint32_t li = p.l_begin + (uint32_t)(pair / n_r), ri = p.r_begin + (uint32_t)(pair % n_r);
        const uint32_t lg0 = __ldg(p.L.item_level_off + li), kl = __ldg(p.L.item_level_off + li + 1) - lg0;
        const uint32_t rg0 = __ldg(p.R.item_level_off + ri), kr = __ldg(p.R.item_level_off + ri + 1) - rg0;
        bool ok = keep_categories(p.job.cat_mode, p.job.cat_mode ? __ldg(p.job.l_cat + li) : 0,
                                  p.job.cat_mode ? __ldg(p.job.r_cat + ri) : 0);
        double score = 0.0;
        if (ok && (kl == 0) != (kr == 0)) {   // IndexError in the reference
            if (lane == 0) atomicOr(p.job.out_flags, NSM_FLAG_EMPTY_ITEM);
            ok = false;
        }
        if (ok && kl && kr) {
            const uint32_t kmax = flat ? 1u : max(kl, kr);
            double w = flat ? 2.0 : 1.0;
            for (uint32_t t = 1; t <= kmax; ++t) {
                const uint32_t gl = lg0 + (flat ? 0u : min(t, kl - 1)), gr = rg0 + (flat ? 0u : min(t, kr - 1));
                const uint32_t na = __ldg(p.L.level_len + gl), nb = __ldg(p.R.level_len + gr);
                uint32_t lcs = 0;
                if (na && nb)
                    lcs = lcs_warp(p.L.chr + __ldg(p.L.level_chr_off + gl), na,
                                   p.R.chr + __ldg(p.R.level_chr_off + gr), nb);
                ++st_evals;
                w *= 0.5;
                score = __fma_rn(qratio_from_lcs(na, nb, lcs), w, score);
            }
        }
        emit_pairs(lane == 0 && ok && score >= thr, p.swap_out ? ri : li, p.swap_out ? li : ri, score, out,
                   p.job.out_capacity, count, p.job.out_flags);
    }
    if (p.job.out_stats && lane == 0 && st_evals) {
        unsigned long long *st = reinterpret_cast<unsigned long long *>(p.job.out_stats);
        atomicAdd(st + NSM_STAT_LEVEL_EVALS, st_evals);
        atomicAdd(st + NSM_STAT_CANDIDATES, st_evals);
    }
}

// left items [l_begin, l_end) x right items [r_begin, r_end), any string lengths
int qratio_long_launch(const nsm_strings_t *left, const nsm_strings_t *right, const nsm_job_t *job,
                       uint32_t l_begin, uint32_t l_end, uint32_t r_begin, uint32_t r_end, bool swap_out,
                       cudaStream_t stream) {
    if (l_end <= l_begin || r_end <= r_begin) return NSM_OK;
    QlongParams p;
    p.L = *left; p.R = *right; p.job = *job;
    p.l_begin = l_begin; p.l_end = l_end; p.r_begin = r_begin; p.r_end = r_end;
    p.swap_out = swap_out ? 1u : 0u;
    const uint32_t shorter = left->max_len < right->max_len ? left->max_len : right->max_len;
    p.words_cap = (shorter + 63u) / 64u;
    if (p.words_cap == 0) p.words_cap = 1;
    const uint32_t n_alpha = left->n_alphabet ? left->n_alphabet : 1u;
    const size_t smem = (size_t)n_alpha * p.words_cap * 8;
    if (p.words_cap > QL_MAX_WORDS || smem > QL_SMEM_BUDGET) {
        set_error("level strings of %u characters over %u codes exceed the generic kernel "
                  "(%u characters, %zu bytes of mask table)", shorter, n_alpha, 64u * QL_MAX_WORDS,
                  QL_SMEM_BUDGET);
        return NSM_ERR_UNSUPPORTED;
    }
    NSM_CUDA_CHECK(cudaFuncSetAttribute(qratio_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    const uint64_t pairs = (uint64_t)(l_end - l_begin) * (r_end - r_begin);
    // as many one-warp CTAs per SM as the mask tables leave room for
    uint64_t per_sm = smem ? (QL_SMEM_BUDGET + 24 * 1024) / (smem + 1024) : 32;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 32) per_sm = 32;
    const uint64_t resident = per_sm * (uint64_t)sm_count();
    const uint32_t grid = (uint32_t)(pairs < resident ? pairs : resident);
    qratio_long_kernel<<<grid, 32, smem, stream>>>(p);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

}  // namespace nsm
