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
// so that the 32 lanes of a warp read 32 consecutive 64-bit words (no bank conflicts).  The left
// tile's level strings are staged in shared memory and used as the text: all threads of the CTA
// consume the same character at the same time, so the inner loop has a uniform trip count and no
// divergence.  Strings longer than 64 characters use W words per thread with the add carry
// chained through registers (W is a template parameter, <= 8, i.e. 512 characters).
//
// compare_terms' level schedule runs as the outer loop (step t uses level min(t, K-1) on both
// sides, weight 2^-t), so a thread rebuilds its masks at most K_right times per tile; partial
// scores of the tile's pairs live in shared memory as float64 and are accumulated in the
// reference's order.  Pairs with score >= threshold are compacted with one atomic per warp.
#include "nsm_common.cuh"

namespace nsm {

constexpr int Q_MAX_THREADS = 256;
constexpr int Q_TILE_LEFT = 32;       // left items per tile (upper bound)
constexpr int Q_CHR_CAP = 24 * 1024;  // bytes of left level strings staged per tile
constexpr int Q_LEV_CAP = 2048;       // left levels staged per tile
constexpr int Q_MAX_WORDS = 8;
constexpr size_t Q_SMEM_BUDGET = 200 * 1024;

struct QratioParams {
    nsm_strings_t L, R;
    nsm_job_t job;
    uint32_t tile_left, n_ltiles, n_rtiles;
    uint32_t threads;   // right items per tile
    uint32_t n_alpha;   // rows of the mask table
};

struct QratioLayout {  // offsets into dynamic shared memory
    size_t pm, acc, chr, lev_off, item_g0, cat, misc, total;
};

__host__ __device__ inline QratioLayout qratio_layout(uint32_t n_alpha, uint32_t words,
                                                      uint32_t threads, uint32_t tile_left) {
    QratioLayout l;
    size_t o = 0;
    l.pm = o;      o += (size_t)n_alpha * words * threads * 8;
    l.acc = o;     o += (size_t)tile_left * threads * 8;
    l.chr = o;     o += Q_CHR_CAP;
    l.lev_off = o; o += (Q_LEV_CAP + 1) * 4;
    o = (o + 7) & ~(size_t)7;
    l.cat = o;     o += Q_TILE_LEFT * 8;
    l.item_g0 = o; o += (Q_TILE_LEFT + 1) * 4;
    l.misc = o;    o += 16;
    l.total = (o + 15) & ~(size_t)15;
    return l;
}

template <int W>
__global__ void __launch_bounds__(Q_MAX_THREADS, 1)
qratio_allpairs_kernel(const QratioParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x, nthr = blockDim.x;
    const QratioLayout lay = qratio_layout(p.n_alpha, W, nthr, p.tile_left);
    uint64_t *s_pm = reinterpret_cast<uint64_t *>(smem_raw + lay.pm);      // [c][w][thread]
    double *s_acc = reinterpret_cast<double *>(smem_raw + lay.acc);        // [li][thread]
    uint8_t *s_chr = smem_raw + lay.chr;
    uint32_t *s_lev_off = reinterpret_cast<uint32_t *>(smem_raw + lay.lev_off);
    uint32_t *s_item_g0 = reinterpret_cast<uint32_t *>(smem_raw + lay.item_g0);
    uint64_t *s_cat = reinterpret_cast<uint64_t *>(smem_raw + lay.cat);
    uint32_t *s_misc = reinterpret_cast<uint32_t *>(smem_raw + lay.misc);  // [0] max K right, [1] max K left

    unsigned long long *count = reinterpret_cast<unsigned long long *>(p.job.out_count);
    const bool flat = p.job.flat != 0;
    const double thr = p.job.threshold;
    unsigned long long st_evals = 0;

    const uint32_t n_tiles = p.n_ltiles * p.n_rtiles;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t lt = tile / p.n_rtiles, rt = tile - lt * p.n_rtiles;
        const uint32_t l0 = p.job.l_row_begin + lt * p.tile_left;
        const uint32_t nl = min(p.tile_left, p.job.l_row_end - l0);
        const uint32_t G0 = __ldg(p.L.item_level_off + l0);
        const uint32_t nlev = __ldg(p.L.item_level_off + l0 + nl) - G0;
        const uint32_t C0 = __ldg(p.L.level_chr_off + G0);
        const uint32_t nchr = __ldg(p.L.level_chr_off + G0 + nlev) - C0;

        __syncthreads();  // previous tile fully consumed
        for (uint32_t i = tid; i < nchr; i += nthr) s_chr[i] = __ldg(p.L.chr + C0 + i);
        for (uint32_t g = tid; g <= nlev; g += nthr) s_lev_off[g] = __ldg(p.L.level_chr_off + G0 + g) - C0;
        if (tid <= nl) s_item_g0[tid] = __ldg(p.L.item_level_off + l0 + tid) - G0;
        if (tid < nl) s_cat[tid] = p.job.cat_mode ? __ldg(p.job.l_cat + l0 + tid) : 0;
        if (tid < 2) s_misc[tid] = 0;
        __syncthreads();

        const uint32_t r = rt * nthr + tid;
        const bool r_valid = r < p.R.n_items;
        uint32_t rg0 = 0, kr = 0;
        uint64_t rcat = 0;
        if (r_valid) {
            rg0 = __ldg(p.R.item_level_off + r);
            kr = __ldg(p.R.item_level_off + r + 1) - rg0;
            if (p.job.cat_mode) rcat = __ldg(p.job.r_cat + r);
        }
        {   // tile-wide level counts bound the schedule
            uint32_t k = kr;
            for (int o = 16; o; o >>= 1) k = max(k, __shfl_xor_sync(FULL_MASK, k, o));
            if ((tid & 31u) == 0) atomicMax(&s_misc[0], k);
            if (tid < nl) atomicMax(&s_misc[1], s_item_g0[tid + 1] - s_item_g0[tid]);
        }
        for (uint32_t li = 0; li < nl; ++li) s_acc[li * nthr + tid] = 0.0;
        __syncthreads();
        const uint32_t max_kr = s_misc[0], max_kl = s_misc[1];
        const uint32_t t_end = flat ? 1u : max(max_kr, max_kl);

        uint32_t cur_slot = 0xffffffffu, m = 0;
        double w = flat ? 2.0 : 1.0;
        for (uint32_t t = 1; t <= t_end; ++t) {
            w *= 0.5;
            if (kr) {
                const uint32_t slot = flat ? 0u : min(t, kr - 1);
                if (slot != cur_slot) {  // (re)build my pattern masks for this right level
                    cur_slot = slot;
                    const uint32_t c0 = __ldg(p.R.level_chr_off + rg0 + slot);
                    m = __ldg(p.R.level_chr_off + rg0 + slot + 1) - c0;
                    for (uint32_t row = 0; row < p.n_alpha * W; ++row) s_pm[row * nthr + tid] = 0;
                    for (uint32_t j = 0; j < m; ++j) {
                        const uint32_t c = __ldg(p.R.chr + c0 + j);
                        s_pm[(c * W + (j >> 6)) * nthr + tid] |= 1ull << (j & 63u);
                    }
                }
            }
            for (uint32_t li = 0; li < nl; ++li) {
                const uint32_t lg0 = s_item_g0[li], kl = s_item_g0[li + 1] - lg0;
                if (kl == 0 || t > max(kl, max_kr)) continue;  // uniform: nobody needs this step
                const uint32_t gl = lg0 + (flat ? 0u : min(t, kl - 1));
                const uint32_t tb = s_lev_off[gl], n = s_lev_off[gl + 1] - tb;
                uint64_t S[W];
#pragma unroll
                for (int x = 0; x < W; ++x) S[x] = ~0ull;
                const uint64_t *pm = s_pm + tid;
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t c = s_chr[tb + j];
                    const uint64_t *row = pm + (size_t)c * W * nthr;
                    uint32_t carry = 0;
#pragma unroll
                    for (int x = 0; x < W; ++x) {
                        const uint64_t M = row[(size_t)x * nthr];
                        const uint64_t u = S[x] & M;
                        const uint64_t sum = S[x] + u;
                        const uint64_t sum2 = sum + carry;
                        if (W > 1) carry = (sum < u) | (sum2 < sum);
                        S[x] = sum2 | (S[x] - u);
                    }
                }
                const bool active = r_valid && kr && (flat || t <= max(kl, kr));
                if (active) {
                    uint32_t lcs = 0;
#pragma unroll
                    for (int x = 0; x < W; ++x) lcs += __popcll(~S[x]);
                    double ratio = 0.0;  // QRatio is 0 when either processed string is empty
                    if (m && n) {
                        const uint32_t lensum = m + n, dist = lensum - 2u * lcs;
                        const double norm_dist = __ddiv_rn((double)dist, (double)lensum);
                        const double norm_sim = __dsub_rn(1.0, norm_dist);
                        ratio = __ddiv_rn(__dmul_rn(norm_sim, 100.0), 100.0);
                    }
                    double *a = s_acc + li * nthr + tid;
                    *a = __fma_rn(ratio, w, *a);
                    ++st_evals;
                }
            }
        }

        for (uint32_t li = 0; li < nl; ++li) {
            const uint32_t kl = s_item_g0[li + 1] - s_item_g0[li];
            bool ok = r_valid && keep_categories(p.job.cat_mode, s_cat[li], rcat);
            if (ok && (kl == 0) != (kr == 0)) {  // IndexError in the reference
                atomicOr(p.job.out_flags, NSM_FLAG_EMPTY_ITEM);
                ok = false;
            }
            const double score = s_acc[li * nthr + tid];
            emit_pairs(ok && score >= thr, l0 + li, r, score, p.job.out_pairs, p.job.out_capacity,
                       count, p.job.out_flags);
        }
    }
    if (p.job.out_stats && st_evals) {
        for (int o = 16; o; o >>= 1) st_evals += __shfl_xor_sync(FULL_MASK, st_evals, o);
        if ((tid & 31u) == 0)
            atomicAdd(reinterpret_cast<unsigned long long *>(p.job.out_stats) + NSM_STAT_LEVEL_EVALS,
                      st_evals);
    }
}

template <int W>
static int launch_qratio(const QratioParams &p, size_t smem, uint32_t grid, cudaStream_t stream) {
    NSM_CUDA_CHECK(cudaFuncSetAttribute(qratio_allpairs_kernel<W>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qratio_allpairs_kernel<W><<<grid, p.threads, smem, stream>>>(p);
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

}  // namespace nsm

extern "C" int nsm_qratio_allpairs(const nsm_strings_t *left, const nsm_strings_t *right,
                                   const nsm_job_t *job, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!left || !right || !job) { set_error("null argument"); return NSM_ERR_BAD_ARG; }
    if (int rc = prepare_job(job, left->n_items, stream)) return rc;
    if (job->l_row_begin == job->l_row_end || right->n_items == 0) return NSM_OK;
    if (job->flat && (left->max_levels > 1 || right->max_levels > 1)) {
        set_error("flat scoring needs items with exactly one level");
        return NSM_ERR_BAD_ARG;
    }
    if (left->n_alphabet != right->n_alphabet || left->n_alphabet > 255) {
        set_error("both sides must be packed with one alphabet of <= 255 codes");
        return NSM_ERR_BAD_ARG;
    }
    const uint32_t words = (right->max_len + 63) / 64 ? (right->max_len + 63) / 64 : 1;
    if (words > (uint32_t)Q_MAX_WORDS) {
        set_error("right level strings of up to %u characters; the kernel handles %d",
                  right->max_len, 64 * Q_MAX_WORDS);
        return NSM_ERR_UNSUPPORTED;
    }
    const uint32_t kl = left->max_levels ? left->max_levels : 1u;
    const uint32_t per_item_chr = kl * (left->max_len ? left->max_len : 1u);
    uint32_t tl = (uint32_t)Q_TILE_LEFT;
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
    // words actually instantiated: 1, 2, 3, 4, 6, 8
    const uint32_t w_inst = words <= 4 ? words : (words <= 6 ? 6u : 8u);
    uint32_t threads = Q_MAX_THREADS;
    while (threads >= 32 && qratio_layout(p.n_alpha, w_inst, threads, tl).total > Q_SMEM_BUDGET)
        threads -= 32;
    if (threads < 32) {
        set_error("alphabet %u x %u words does not fit shared memory", p.n_alpha, w_inst);
        return NSM_ERR_UNSUPPORTED;
    }
    p.threads = threads;
    const size_t smem = qratio_layout(p.n_alpha, w_inst, threads, tl).total;
    const uint32_t n_rows = job->l_row_end - job->l_row_begin;
    p.n_ltiles = (n_rows + tl - 1) / tl;
    p.n_rtiles = (right->n_items + threads - 1) / threads;
    const uint64_t n_tiles64 = (uint64_t)p.n_ltiles * p.n_rtiles;
    if (n_tiles64 > 0xffffffffull) {
        set_error("too many tiles (%llu); split the left row block", (unsigned long long)n_tiles64);
        return NSM_ERR_UNSUPPORTED;
    }
    const uint32_t resident = (uint32_t)sm_count();
    const uint32_t grid = (uint32_t)(n_tiles64 < resident ? n_tiles64 : resident);
    switch (w_inst) {
        case 1: return launch_qratio<1>(p, smem, grid, stream);
        case 2: return launch_qratio<2>(p, smem, grid, stream);
        case 3: return launch_qratio<3>(p, smem, grid, stream);
        case 4: return launch_qratio<4>(p, smem, grid, stream);
        case 6: return launch_qratio<6>(p, smem, grid, stream);
        default: return launch_qratio<8>(p, smem, grid, stream);
    }
}
