// intersection_vs_union all-pairs kernel (sm_100a).
//
// Replaces the pair loop of ComparableData.gen_comparable for score_func ==
// "intersection_vs_union" (/root/reference/napkon_string_matching/types/comparable_data.py:223-243,
// compare_terms :248-265, compare/score_functions.py:6-13).
//
// Work decomposition: a tile is <= JT_LEFT left items x JT_THREADS right items.  Each thread owns
// one right item; the left tile's level records (signature, size word, token offset) are staged
// in shared memory and read by broadcast while the CTA walks the left items.  Per item pair:
//   1. FILTER (integer + fp32, round-up): upper-bound every level score from the 64-bit token
//      signatures, I <= min(popc(sigL & sigR) + min(exL, exR), |A|, |B|), accumulate the
//      compare_terms weights with directed rounding and stop as soon as bound + remaining
//      weight cannot reach the threshold.  A pair that fails the filter is proven < threshold.
//   2. EXACT: pairs that pass are compacted per warp (ballot + popc into a shared queue) so that
//      32 lanes score 32 surviving pairs: per used level the exact |A & B| (popc when the
//      signature is exact or empty, else a merge over the sorted ids), the same int/int float64
//      division and the same accumulation order as the reference.
//   3. COMPACTION: pairs with score >= threshold are appended to the output with one atomic
//      per warp (nsm_common.cuh: emit_pairs).
#include "nsm_common.cuh"

namespace nsm {

constexpr int JT_THREADS = 256;  // right items per tile (= threads per CTA)
constexpr int JT_LEFT = 64;      // left items per tile (upper bound)
constexpr int J_LCAP = 4096;     // left levels staged per tile (upper bound)
constexpr int J_RCP = 512;       // reciprocal table size
constexpr int J_WARPS = JT_THREADS / 32;

struct JaccardParams {
    nsm_sets_t L, R;
    nsm_job_t job;
    float thr_lo;       // filter threshold (see filter_threshold)
    uint32_t tile_left; // left items per tile, tile_left * L.max_levels <= J_LCAP
    uint32_t n_ltiles, n_rtiles;
};

struct __align__(16) JaccardSmem {
    uint64_t sig[J_LCAP];
    uint32_t info[J_LCAP];
    uint32_t tok_off[J_LCAP + 1];
    uint32_t item_g0[JT_LEFT + 1];  // tile-relative first level of each left item
    uint64_t cat[JT_LEFT];
    float rcp_up[J_RCP];
    uint32_t queue[J_WARPS][64];
    unsigned long long stats[NSM_N_STATS];
};

__device__ __forceinline__ float pow2_neg(uint32_t k) {  // 2^-k, 0 when it underflows fp32
    return k <= 126 ? __int_as_float((127 - (int)k) << 23) : 0.0f;
}

__device__ __forceinline__ uint32_t merge_count(const uint32_t *__restrict__ a, uint32_t na,
                                                const uint32_t *__restrict__ b, uint32_t nb) {
    if (na == 0 || nb == 0) return 0;
    uint32_t i = 0, j = 0, c = 0;
    uint32_t x = __ldg(a), y = __ldg(b);
    while (true) {
        if (x == y) {
            ++c; ++i; ++j;
            if (i >= na || j >= nb) break;
            x = __ldg(a + i); y = __ldg(b + j);
        } else if (x < y) {
            if (++i >= na) break;
            x = __ldg(a + i);
        } else {
            if (++j >= nb) break;
            y = __ldg(b + j);
        }
    }
    return c;
}

__global__ void __launch_bounds__(JT_THREADS, 2)
jaccard_allpairs_kernel(const JaccardParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    JaccardSmem &s = *reinterpret_cast<JaccardSmem *>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long *count = reinterpret_cast<unsigned long long *>(p.job.out_count);
    const bool flat = p.job.flat != 0;
    const bool sig_exact = p.L.sig_exact != 0 && p.R.sig_exact != 0;
    const double thr = p.job.threshold;

    for (unsigned u = tid; u < J_RCP; u += JT_THREADS) s.rcp_up[u] = u ? __frcp_ru((float)u) : 0.0f;
    if (tid < NSM_N_STATS) s.stats[tid] = 0;
    unsigned long long st_cand = 0, st_evals = 0, st_merges = 0;

    const uint32_t n_tiles = p.n_ltiles * p.n_rtiles;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t lt = tile / p.n_rtiles, rt = tile - lt * p.n_rtiles;
        const uint32_t l0 = p.job.l_row_begin + lt * p.tile_left;
        const uint32_t nl = min(p.tile_left, p.job.l_row_end - l0);
        const uint32_t G0 = __ldg(p.L.item_level_off + l0);
        const uint32_t nlev = __ldg(p.L.item_level_off + l0 + nl) - G0;

        __syncthreads();  // previous tile fully consumed
        for (uint32_t g = tid; g < nlev; g += JT_THREADS) {
            s.sig[g] = __ldg(p.L.level_sig + G0 + g);
            s.info[g] = __ldg(p.L.level_info + G0 + g);
            s.tok_off[g] = __ldg(p.L.level_tok_off + G0 + g);
        }
        if (tid == 0) s.tok_off[nlev] = __ldg(p.L.level_tok_off + G0 + nlev);
        if (tid <= nl) s.item_g0[tid] = __ldg(p.L.item_level_off + l0 + tid) - G0;
        if (tid < nl) s.cat[tid] = p.job.cat_mode ? __ldg(p.job.l_cat + l0 + tid) : 0;
        __syncthreads();

        // my right item
        const uint32_t r = rt * JT_THREADS + tid;
        const bool r_valid = r < p.R.n_items;
        uint32_t rg0 = 0, kr = 0;
        uint64_t rcat = 0;
        if (r_valid) {
            rg0 = __ldg(p.R.item_level_off + r);
            kr = __ldg(p.R.item_level_off + r + 1) - rg0;
            if (p.job.cat_mode) rcat = __ldg(p.job.r_cat + r);
        }
        // slot 1 (the level compare_terms weights with 1/2) is needed for every left item
        uint64_t rsig1 = 0;
        uint32_t rinfo1 = 0;
        if (r_valid && kr) {
            const uint32_t g = rg0 + (flat ? 0u : min(1u, kr - 1));
            rsig1 = __ldg(p.R.level_sig + g);
            rinfo1 = __ldg(p.R.level_info + g);
        }

        uint32_t qn = 0;  // warp-uniform queue fill
        auto score_candidate = [&](bool active, uint32_t entry) {
            // entry = left item (tile-relative) << 5 | lane owning the right item
            const uint32_t li = entry >> 5, rl = entry & 31u;
            const uint32_t c_rg0 = __shfl_sync(FULL_MASK, rg0, rl);
            const uint32_t c_kr = __shfl_sync(FULL_MASK, kr, rl);
            const uint32_t c_r = rt * JT_THREADS + (warp << 5) + rl;
            double score = 0.0;
            bool ok = active;
            if (active) {
                const uint32_t lg0 = s.item_g0[li], kl = s.item_g0[li + 1] - lg0;
                const uint32_t kmax = flat ? 1u : max(kl, c_kr);
                ++st_cand;
                if (kl == 0 || c_kr == 0) {
                    // both empty: compare_terms returns 0; one empty: IndexError in the reference
                    if (kl != c_kr) { atomicOr(p.job.out_flags, NSM_FLAG_EMPTY_ITEM); ok = false; }
                } else {
                    double w = flat ? 2.0 : 1.0;
                    uint32_t pgl = 0xffffffffu, pgr = 0xffffffffu, inter = 0, uni = 1;
                    for (uint32_t t = 1; t <= kmax; ++t) {
                        const uint32_t gl = lg0 + (flat ? 0u : min(t, kl - 1));
                        const uint32_t gr = c_rg0 + (flat ? 0u : min(t, c_kr - 1));
                        if (gl != pgl || gr != pgr) {
                            pgl = gl; pgr = gr;
                            const uint64_t sl = s.sig[gl], sr = __ldg(p.R.level_sig + gr);
                            const uint32_t a = s.info[gl] & 0xffffu;
                            const uint32_t b = __ldg(p.R.level_info + gr) & 0xffffu;
                            const uint64_t both = sl & sr;
                            ++st_evals;
                            if (both == 0) {
                                inter = 0;
                            } else if (sig_exact) {
                                inter = __popcll(both);
                            } else {
                                const uint32_t ta = s.tok_off[gl], tb = __ldg(p.R.level_tok_off + gr);
                                inter = merge_count(p.L.tok + ta, a, p.R.tok + tb, b);
                                ++st_merges;
                            }
                            uni = a + b - inter;
                        } else {
                            ++st_evals;
                        }
                        w *= 0.5;
                        // len(A & B) / len(A | B): int / int true division, then score += s * w
                        const double sc = __ddiv_rn((double)inter, (double)uni);
                        if (uni == 0) { atomicOr(p.job.out_flags, NSM_FLAG_ZERO_UNION); ok = false; }
                        score = __fma_rn(sc, w, score);
                    }
                }
            }
            const uint32_t c_l = l0 + li;
            emit_pairs(ok && score >= thr, c_l, c_r, score, p.job.out_pairs, p.job.out_capacity,
                       count, p.job.out_flags);
        };

        for (uint32_t li = 0; li < nl; ++li) {
            const uint32_t lg0 = s.item_g0[li], kl = s.item_g0[li + 1] - lg0;
            bool pass = r_valid && keep_categories(p.job.cat_mode, s.cat[li], rcat);
            if (pass) {
                if (kl == 0 || kr == 0) {
                    pass = true;  // rare; the exact path sorts out 0 vs IndexError
                } else {
                    const uint32_t kmax = flat ? 1u : max(kl, kr);
                    const float w_last = flat ? 1.0f : pow2_neg(kmax);
                    float w = flat ? 2.0f : 1.0f, ub = 0.0f;
                    for (uint32_t t = 1; t <= kmax; ++t) {
                        const uint32_t gl = lg0 + (flat ? 0u : min(t, kl - 1));
                        uint64_t sr;
                        uint32_t ir;
                        if (t == 1) {
                            sr = rsig1; ir = rinfo1;
                        } else {
                            const uint32_t gr = rg0 + min(t, kr - 1);
                            sr = __ldg(p.R.level_sig + gr);
                            ir = __ldg(p.R.level_info + gr);
                        }
                        const uint64_t sl = s.sig[gl];
                        const uint32_t il = s.info[gl];
                        const uint32_t a = il & 0xffffu, b = ir & 0xffffu;
                        const uint32_t ex = min((il >> 16) & 0xffu, (ir >> 16) & 0xffu);
                        // no shared bit -> no shared token; else |A & B| <= shared bits + the
                        // tokens either side folded onto an occupied bit (255 = saturated count)
                        uint32_t ih = __popcll(sl & sr);
                        ih = ih ? min(min(ex == 255u ? 0xffffu : ih + ex, a), b) : 0u;
                        const uint32_t uh = a + b - ih;
                        w = fmaxf(w * 0.5f, 1.17549435e-38f);
                        if (ih) {
                            const float rc = uh < J_RCP ? s.rcp_up[uh] : __frcp_ru((float)uh);
                            ub = __fmaf_ru(__fmul_ru((float)ih, rc), w, ub);
                        }
                        // weights still to come: 2^-t - 2^-kmax
                        const float rem = __fsub_ru(w, w_last);
                        if (__fadd_ru(ub, rem) < p.thr_lo) { pass = false; break; }
                    }
                    if (pass) pass = ub >= p.thr_lo;
                }
            }
            const unsigned m = __ballot_sync(FULL_MASK, pass);
            if (m) {
                if (pass) s.queue[warp][qn + __popc(m & lanemask_lt())] = (li << 5) | lane;
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    const uint32_t entry = s.queue[warp][qn + lane];
                    __syncwarp();
                    score_candidate(true, entry);
                }
            }
        }
        if (qn) {
            const bool active = lane < qn;
            const uint32_t entry = active ? s.queue[warp][lane] : 0u;
            __syncwarp();
            score_candidate(active, entry);
        }
    }

    if (p.job.out_stats) {
        atomicAdd(&s.stats[NSM_STAT_CANDIDATES], st_cand);
        atomicAdd(&s.stats[NSM_STAT_LEVEL_EVALS], st_evals);
        atomicAdd(&s.stats[NSM_STAT_LEVEL_MERGES], st_merges);
        __syncthreads();
        if (tid < NSM_N_STATS && s.stats[tid])
            atomicAdd(reinterpret_cast<unsigned long long *>(p.job.out_stats) + tid, s.stats[tid]);
    }
}

}  // namespace nsm

extern "C" int nsm_jaccard_allpairs(const nsm_sets_t *left, const nsm_sets_t *right,
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
    if (left->max_levels > (uint32_t)J_LCAP) {
        set_error("left items have up to %u levels; the kernel stages at most %d", left->max_levels,
                  J_LCAP);
        return NSM_ERR_UNSUPPORTED;
    }

    JaccardParams p;
    p.L = *left; p.R = *right; p.job = *job;
    p.thr_lo = filter_threshold(job->threshold);
    const uint32_t kl = left->max_levels ? left->max_levels : 1u;
    uint32_t tl = (uint32_t)J_LCAP / kl;
    p.tile_left = tl < 1 ? 1u : (tl > (uint32_t)JT_LEFT ? (uint32_t)JT_LEFT : tl);
    const uint32_t n_rows = job->l_row_end - job->l_row_begin;
    p.n_ltiles = (n_rows + p.tile_left - 1) / p.tile_left;
    p.n_rtiles = (right->n_items + JT_THREADS - 1) / JT_THREADS;
    const uint64_t n_tiles64 = (uint64_t)p.n_ltiles * p.n_rtiles;
    if (n_tiles64 > 0xffffffffull) {
        set_error("too many tiles (%llu); split the left row block", (unsigned long long)n_tiles64);
        return NSM_ERR_UNSUPPORTED;
    }

    static bool attr_set = false;
    const size_t smem = sizeof(JaccardSmem);
    if (!attr_set) {
        NSM_CUDA_CHECK(cudaFuncSetAttribute(jaccard_allpairs_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const uint32_t resident = 2u * (uint32_t)sm_count();  // __launch_bounds__(.., 2)
    const uint32_t grid = (uint32_t)(n_tiles64 < resident ? n_tiles64 : resident);
    jaccard_allpairs_kernel<<<grid, JT_THREADS, smem, stream>>>(p);
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}
