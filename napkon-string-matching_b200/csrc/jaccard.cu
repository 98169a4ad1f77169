// intersection_vs_union all-pairs kernel (sm_100a).
//
// Replaces the pair loop of ComparableData.gen_comparable for score_func ==
// "intersection_vs_union" (/root/reference/napkon_string_matching/types/comparable_data.py:223-243,
// compare_terms :248-265, compare/score_functions.py:6-13).
//
// Data: every level set carries a 128-bit summary (pack.py): `head` = exact bitset of its 64 most
// frequent vocabulary ids, `tail` = signature of the rest, plus sizes.  Every item carries the OR
// of those over the levels compare_terms can touch (`item_any`).
//
// Work decomposition: a unit is one block of JT_THREADS right items x a group of left tiles
// (<= JT_LEFT items each).  The right block is staged once per unit in shared memory, slot-major
// (slot t = the level compare_terms uses at step t, i.e. min(t, K-1)) so that lane i reads word i:
// conflict-free.  The left tile is staged as CSR and read by broadcast.  Each warp then runs a
// three-stage funnel over its 32 right items x the tile's left items, re-compacting the survivors
// of every stage through a per-warp shared-memory queue (ballot + popc) so that each stage runs
// with 32 busy lanes:
//   A  ANY      (per item pair, ~6 integer ops) item_any(left) & item_any(right) == 0 proves that
//               no level pair shares a token: score 0, below any positive threshold.
//   B  BOUND    (per level, integer + fp32 with round-up) exact head intersection by popcount plus
//               an upper bound of the tail intersection from the signatures gives an upper bound
//               of every level score; the compare_terms weights are accumulated with directed
//               rounding, stopping when bound + remaining weight cannot reach the threshold.
//   C  EXACT    per used level the exact |A & B| (popcount; a merge over the tail ids only when
//               the tail signatures collide), the reference's int/int float64 division and its
//               accumulation order; score >= threshold in float64.
// Kept pairs go to a per-warp staging buffer and are flushed with one global atomic per ~100
// records as coalesced 16-byte stores.
#include "nsm_common.cuh"

namespace nsm {

constexpr int JT_THREADS = 256;  // right items per block (= threads per CTA)
constexpr int JT_LEFT = 64;      // left items per tile (upper bound)
constexpr int J_LCAP = 1024;     // left levels staged per tile (upper bound)
constexpr int J_RSLOTS = 10;     // right slots staged per item; deeper steps gather from global
constexpr int J_GROUP = 8;       // left tiles per unit
constexpr int J_RCP = 512;       // reciprocal table size
constexpr int J_WARPS = JT_THREADS / 32;
constexpr int J_OUT = 128;       // staged output records per warp

struct JaccardParams {
    nsm_sets_t L, R;
    nsm_job_t job;
    float thr_lo;        // filter threshold (see filter_threshold); -inf: everything passes
    uint32_t tile_left;  // left items per tile, tile_left * L.max_levels <= J_LCAP
    uint32_t n_ltiles, n_lgroups, n_rblocks;
};

struct __align__(16) JaccardSmem {
    // left tile, CSR
    uint64_t l_head[J_LCAP];
    uint64_t l_tail[J_LCAP];
    uint32_t l_info[J_LCAP];
    uint32_t l_tok_off[J_LCAP + 1];
    uint32_t l_g0[JT_LEFT + 1];  // tile-relative first level of each left item
    ulonglong2 l_any[JT_LEFT];
    uint64_t l_cat[JT_LEFT];
    // right block, slot-major
    uint64_t r_head[J_RSLOTS][JT_THREADS];
    uint64_t r_tail[J_RSLOTS][JT_THREADS];
    uint32_t r_info[J_RSLOTS][JT_THREADS];
    uint32_t r_g0[JT_THREADS];
    uint32_t r_k[JT_THREADS];
    float rcp_up[J_RCP];
    uint32_t qa[J_WARPS][64];
    uint32_t qb[J_WARPS][64];
    nsm_pair_t out[J_WARPS][J_OUT];
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

struct LevelWords {
    uint64_t head, tail;
    uint32_t info;
};

__global__ void __launch_bounds__(JT_THREADS, 2)
jaccard_allpairs_kernel(const JaccardParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    JaccardSmem &s = *reinterpret_cast<JaccardSmem *>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long *count = reinterpret_cast<unsigned long long *>(p.job.out_count);
    const bool flat = p.job.flat != 0;
    const bool exact_bits = p.L.exact_bits != 0 && p.R.exact_bits != 0;
    const double thr = p.job.threshold;
    const bool pass_all = !(p.thr_lo > -INFINITY) && !(p.thr_lo != p.thr_lo);  // thr <= 0

    for (unsigned u = tid; u < J_RCP; u += JT_THREADS) s.rcp_up[u] = u ? __frcp_ru((float)u) : 0.0f;
    if (tid < NSM_N_STATS) s.stats[tid] = 0;
    unsigned long long st_cand = 0, st_evals = 0, st_merges = 0, st_bound = 0;

    uint32_t out_n = 0;  // warp-uniform fill of s.out[warp]
    auto flush_out = [&]() {
        if (out_n == 0) return;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(count, (unsigned long long)out_n);
        base = __shfl_sync(FULL_MASK, base, 0);
        const double2 *src = reinterpret_cast<const double2 *>(&s.out[warp][0]);
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
    auto emit = [&](bool keep, uint32_t left, uint32_t right, double score) {  // all 32 lanes
        const unsigned m = __ballot_sync(FULL_MASK, keep);
        if (m == 0) return;
        const uint32_t n = __popc(m);
        if (out_n + n > J_OUT) flush_out();
        if (keep) {
            double2 rec;
            rec.x = __longlong_as_double((long long)(((unsigned long long)right << 32) | left));
            rec.y = score;
            reinterpret_cast<double2 *>(&s.out[warp][0])[out_n + __popc(m & lanemask_lt())] = rec;
        }
        out_n += n;
        __syncwarp();
    };

    const uint32_t n_units = p.n_lgroups * p.n_rblocks;
    for (uint32_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const uint32_t rb = unit / p.n_lgroups, lgroup = unit - rb * p.n_lgroups;

        __syncthreads();  // previous unit fully consumed
        // ---- stage my right item, slot-major --------------------------------------------
        const uint32_t r = rb * JT_THREADS + tid;
        const bool r_valid = r < p.R.n_items;
        uint32_t rg0 = 0, kr = 0;
        uint64_t rcat = 0, rany_h = 0, rany_t = 0;
        if (r_valid) {
            rg0 = __ldg(p.R.item_level_off + r);
            kr = __ldg(p.R.item_level_off + r + 1) - rg0;
            if (p.job.cat_mode) rcat = __ldg(p.job.r_cat + r);
            const ulonglong2 any = __ldg(reinterpret_cast<const ulonglong2 *>(p.R.item_any) + r);
            rany_h = any.x; rany_t = any.y;
        }
        s.r_g0[tid] = rg0;
        s.r_k[tid] = kr;
#pragma unroll
        for (int sl = 0; sl < J_RSLOTS; ++sl) {
            uint64_t h = 0, t = 0;
            uint32_t inf = 0;
            if (kr) {
                const uint32_t g = rg0 + (flat ? 0u : min((uint32_t)sl + 1u, kr - 1));
                h = __ldg(p.R.level_head + g);
                t = __ldg(p.R.level_tail + g);
                inf = __ldg(p.R.level_info + g);
            }
            s.r_head[sl][tid] = h;
            s.r_tail[sl][tid] = t;
            s.r_info[sl][tid] = inf;
        }

        // level words of the right item in column `rc` at step t
        auto right_level = [&](uint32_t rc, uint32_t t, uint32_t c_rg0, uint32_t c_kr) {
            LevelWords w;
            if (t <= (uint32_t)J_RSLOTS) {
                w.head = s.r_head[t - 1][rc];
                w.tail = s.r_tail[t - 1][rc];
                w.info = s.r_info[t - 1][rc];
            } else {
                const uint32_t g = c_rg0 + min(t, c_kr - 1);
                w.head = __ldg(p.R.level_head + g);
                w.tail = __ldg(p.R.level_tail + g);
                w.info = __ldg(p.R.level_info + g);
            }
            return w;
        };

        const uint32_t lt_begin = lgroup * J_GROUP;
        const uint32_t lt_end = min(lt_begin + (uint32_t)J_GROUP, p.n_ltiles);
        for (uint32_t lt = lt_begin; lt < lt_end; ++lt) {
            const uint32_t l0 = p.job.l_row_begin + lt * p.tile_left;
            const uint32_t nl = min(p.tile_left, p.job.l_row_end - l0);
            const uint32_t G0 = __ldg(p.L.item_level_off + l0);
            const uint32_t nlev = __ldg(p.L.item_level_off + l0 + nl) - G0;

            __syncthreads();  // previous tile consumed (and the right block staged)
            for (uint32_t g = tid; g < nlev; g += JT_THREADS) {
                s.l_head[g] = __ldg(p.L.level_head + G0 + g);
                s.l_tail[g] = __ldg(p.L.level_tail + G0 + g);
                s.l_info[g] = __ldg(p.L.level_info + G0 + g);
                s.l_tok_off[g] = __ldg(p.L.level_tok_off + G0 + g);
            }
            if (tid == 0) s.l_tok_off[nlev] = __ldg(p.L.level_tok_off + G0 + nlev);
            if (tid <= nl) s.l_g0[tid] = __ldg(p.L.item_level_off + l0 + tid) - G0;
            if (tid < nl) {
                s.l_any[tid] = __ldg(reinterpret_cast<const ulonglong2 *>(p.L.item_any) + l0 + tid);
                s.l_cat[tid] = p.job.cat_mode ? __ldg(p.job.l_cat + l0 + tid) : 0;
            }
            __syncthreads();

            uint32_t qa_n = 0, qb_n = 0;  // warp-uniform queue fills

            // ---- stage C: exact score of one candidate per lane ---------------------------
            auto stage_exact = [&](bool active, uint32_t entry) {
                const uint32_t li = entry >> 5, rc = (warp << 5) | (entry & 31u);
                double score = 0.0;
                bool ok = active;
                if (active) {
                    const uint32_t lg0 = s.l_g0[li], kl = s.l_g0[li + 1] - lg0;
                    const uint32_t c_rg0 = s.r_g0[rc], c_kr = s.r_k[rc];
                    ++st_cand;
                    if (kl == 0 || c_kr == 0) {
                        // both empty: compare_terms returns 0; one empty: IndexError in the reference
                        if (kl != c_kr) { atomicOr(p.job.out_flags, NSM_FLAG_EMPTY_ITEM); ok = false; }
                    } else {
                        const uint32_t kmax = flat ? 1u : max(kl, c_kr);
                        double w = flat ? 2.0 : 1.0;
                        uint32_t pgl = 0xffffffffu, pgr = 0xffffffffu, inter = 0, uni = 1;
                        for (uint32_t t = 1; t <= kmax; ++t) {
                            const uint32_t gl = lg0 + (flat ? 0u : min(t, kl - 1));
                            const uint32_t gr = c_rg0 + (flat ? 0u : min(t, c_kr - 1));
                            ++st_evals;
                            if (gl != pgl || gr != pgr) {
                                pgl = gl; pgr = gr;
                                const LevelWords R = right_level(rc, t, c_rg0, c_kr);
                                const uint32_t il = s.l_info[gl];
                                const uint32_t a = il & 0xffffu, b = R.info & 0xffffu;
                                inter = __popcll(s.l_head[gl] & R.head);
                                const uint64_t tb = s.l_tail[gl] & R.tail;
                                if (tb) {
                                    if (exact_bits) {
                                        inter += __popcll(tb);
                                    } else {  // ids are sorted: the tail ids follow the n_head head ids
                                        const uint32_t hl = il >> 24, hr = R.info >> 24;
                                        inter += merge_count(p.L.tok + s.l_tok_off[gl] + hl, a - hl,
                                                             p.R.tok + __ldg(p.R.level_tok_off + gr) + hr,
                                                             b - hr);
                                        ++st_merges;
                                    }
                                }
                                uni = a + b - inter;
                            }
                            w *= 0.5;
                            // len(A & B) / len(A | B): int / int true division, then score += s * w
                            const double sc = __ddiv_rn((double)inter, (double)uni);
                            if (uni == 0) { atomicOr(p.job.out_flags, NSM_FLAG_ZERO_UNION); ok = false; }
                            score = __fma_rn(sc, w, score);
                        }
                    }
                }
                emit(ok && score >= thr, l0 + (entry >> 5), rb * JT_THREADS + rc, score);
            };

            // ---- stage B: fp32 upper bound of one surviving pair per lane -----------------
            auto stage_bound = [&](bool active, uint32_t entry) {
                const uint32_t li = entry >> 5, rc = (warp << 5) | (entry & 31u);
                bool pass = active;
                if (active && !pass_all) {
                    const uint32_t lg0 = s.l_g0[li], kl = s.l_g0[li + 1] - lg0;
                    const uint32_t c_rg0 = s.r_g0[rc], c_kr = s.r_k[rc];
                    ++st_bound;
                    if (kl != 0 && c_kr != 0) {
                        const uint32_t kmax = flat ? 1u : max(kl, c_kr);
                        const float w_last = flat ? 1.0f : pow2_neg(kmax);
                        float w = flat ? 2.0f : 1.0f, ub = 0.0f;
                        for (uint32_t t = 1; t <= kmax; ++t) {
                            const uint32_t gl = lg0 + (flat ? 0u : min(t, kl - 1));
                            const LevelWords R = right_level(rc, t, c_rg0, c_kr);
                            const uint32_t il = s.l_info[gl];
                            const uint32_t a = il & 0xffffu, b = R.info & 0xffffu;
                            uint32_t ih = __popcll(s.l_head[gl] & R.head);
                            const uint64_t tb = s.l_tail[gl] & R.tail;
                            if (tb) {
                                // shared tail bits + the ids either side folded onto an occupied bit
                                // bound the shared tail ids (255 = saturated fold count)
                                const uint32_t ex = min((il >> 16) & 0xffu, (R.info >> 16) & 0xffu);
                                uint32_t it = exact_bits ? __popcll(tb)
                                                         : (ex == 255u ? 0xffffu : __popcll(tb) + ex);
                                it = min(it, min(a - (il >> 24), b - (R.info >> 24)));
                                ih += it;
                            }
                            const uint32_t uh = a + b - ih;
                            w = fmaxf(w * 0.5f, 1.17549435e-38f);
                            if (ih) {
                                const float rc_up = uh < J_RCP ? s.rcp_up[uh] : __frcp_ru((float)uh);
                                ub = __fmaf_ru(__fmul_ru((float)ih, rc_up), w, ub);
                            }
                            // weights still to come: 2^-t - 2^-kmax
                            if (__fadd_ru(ub, __fsub_ru(w, w_last)) < p.thr_lo) { pass = false; break; }
                        }
                        if (pass) pass = ub >= p.thr_lo;
                    }
                }
                const unsigned m = __ballot_sync(FULL_MASK, pass);
                if (m) {
                    if (pass) s.qb[warp][qb_n + __popc(m & lanemask_lt())] = entry;
                    qb_n += __popc(m);
                    __syncwarp();
                    if (qb_n >= 32) {
                        qb_n -= 32;
                        const uint32_t e = s.qb[warp][qb_n + lane];
                        __syncwarp();
                        stage_exact(true, e);
                    }
                }
            };

            // ---- stage A: one left item x my right item ------------------------------------
            for (uint32_t li = 0; li < nl; ++li) {
                const ulonglong2 lany = s.l_any[li];
                bool pass = r_valid && keep_categories(p.job.cat_mode, s.l_cat[li], rcat);
                if (pass && !pass_all) {
                    const bool shared = ((lany.x & rany_h) | (lany.y & rany_t)) != 0;
                    const bool l_empty = s.l_g0[li + 1] == s.l_g0[li];
                    pass = shared || (l_empty != (kr == 0));  // the latter: IndexError upstream
                }
                const unsigned m = __ballot_sync(FULL_MASK, pass);
                if (m) {
                    if (pass) s.qa[warp][qa_n + __popc(m & lanemask_lt())] = (li << 5) | lane;
                    qa_n += __popc(m);
                    __syncwarp();
                    if (qa_n >= 32) {
                        qa_n -= 32;
                        const uint32_t e = s.qa[warp][qa_n + lane];
                        __syncwarp();
                        stage_bound(true, e);
                    }
                }
            }
            // drain what is left of both queues before the tile is replaced
            if (qa_n) {
                const bool active = lane < qa_n;
                const uint32_t e = active ? s.qa[warp][lane] : 0u;
                __syncwarp();
                qa_n = 0;
                stage_bound(active, e);
            }
            if (qb_n) {
                const bool active = lane < qb_n;
                const uint32_t e = active ? s.qb[warp][lane] : 0u;
                __syncwarp();
                qb_n = 0;
                stage_exact(active, e);
            }
        }
    }
    flush_out();

    if (p.job.out_stats) {
        atomicAdd(&s.stats[NSM_STAT_CANDIDATES], st_cand);
        atomicAdd(&s.stats[NSM_STAT_LEVEL_EVALS], st_evals);
        atomicAdd(&s.stats[NSM_STAT_LEVEL_MERGES], st_merges);
        atomicAdd(&s.stats[NSM_STAT_BOUND_PAIRS], st_bound);
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
    p.n_lgroups = (p.n_ltiles + J_GROUP - 1) / J_GROUP;
    p.n_rblocks = (right->n_items + JT_THREADS - 1) / JT_THREADS;
    const uint64_t n_units64 = (uint64_t)p.n_lgroups * p.n_rblocks;
    if (n_units64 > 0xffffffffull) {
        set_error("too many work units (%llu); split the left row block",
                  (unsigned long long)n_units64);
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
    const uint32_t grid = (uint32_t)(n_units64 < resident ? n_units64 : resident);
    jaccard_allpairs_kernel<<<grid, JT_THREADS, smem, stream>>>(p);
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}
