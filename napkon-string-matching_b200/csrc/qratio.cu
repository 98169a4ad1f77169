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
// Items with one level each (the flat score function, config 3's `Question` strings) go through
// qratio_flat.cu, which prunes with a sound distance bound; pairs whose items both hold a level
// string of more than 512 characters through qratio_long.cu (one warp per pair).  This file:
//   LEVELS = true   compare_terms' schedule runs as the outer loop per left tile (step t uses
//                   level min(t, K-1) on both sides, weight 2^-t); a thread rebuilds its masks at
//                   most K_right times per tile; partial scores of the tile's pairs live in
//                   shared memory as float64 and are accumulated in the reference's order.
// Pairs with score >= threshold are compacted with one atomic per warp.
#include "qratio_common.cuh"

namespace nsm {

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

    // units are drawn from a device counter, in falling order of cost (both sides are stored by
    // rising length), as in qratio_flat.cu
    const uint32_t n_units = p.n_lgroups * p.n_rblocks;
    uint32_t unit = blockIdx.x;
    while (unit < n_units) {
        const uint32_t unit_rev = n_units - 1u - unit;
        const uint32_t rb = unit_rev / p.n_lgroups, lgroup = unit_rev - rb * p.n_lgroups;
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
        if (LEVELS) {
            __syncthreads();  // s_misc of the previous unit consumed
            if (tid == 0) {
                s_misc[0] = 0;
                s_misc[1] = p.unit_counter ? gridDim.x + atomicAdd(p.unit_counter, 1u) : unit + gridDim.x;
            }
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

            {
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
                    emit_pairs(ok && score >= thr, p.swap_out ? r : l0 + li, p.swap_out ? l0 + li : r, score,
                               static_cast<nsm_pair_t *>(p.job.out_pairs), p.job.out_capacity, count,
                               p.job.out_flags);
                }
            }
        }
        unit = LEVELS ? s_misc[1] : unit + gridDim.x;   // drawn at the top of this unit (see there)
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

static int dispatch_levels(uint32_t words, const QratioParams &p, size_t smem, uint32_t grid,
                           cudaStream_t stream) {
    switch (words) {
        case 1: return launch_qratio<1, true>(p, smem, grid, stream);
        case 2: return launch_qratio<2, true>(p, smem, grid, stream);
        case 3: return launch_qratio<3, true>(p, smem, grid, stream);
        case 4: return launch_qratio<4, true>(p, smem, grid, stream);
        case 6: return launch_qratio<6, true>(p, smem, grid, stream);
        default: return launch_qratio<8, true>(p, smem, grid, stream);
    }
}

// compare_terms over levels: the job's left rows x the right items [r_lo, r_hi) of the classes
// <= Q_MAX_WORDS words.  left_max_len: longest level string among the job's left rows (sizes the
// tile).  Returns NSM_ERR_UNSUPPORTED (without an error text) when one left item does not fit the
// staged tile; the caller then takes the generic kernel for those rows.
static int levels_launch(const nsm_strings_t *left, const nsm_strings_t *right, const nsm_job_t *job,
                         uint32_t left_max_len, uint32_t r_lo, uint32_t r_hi, bool swap_out,
                         cudaStream_t stream) {
    if (job->l_row_end <= job->l_row_begin || r_hi <= r_lo) return NSM_OK;
    const uint32_t kl = left->max_levels ? left->max_levels : 1u;
    const uint32_t per_item_chr = kl * (((left_max_len ? left_max_len : 1u) + 7u) & ~7u);
    uint32_t tl = (uint32_t)Q_TILE_LEVELS;
    if (per_item_chr * tl > (uint32_t)Q_CHR_CAP) tl = (uint32_t)Q_CHR_CAP / per_item_chr;
    if (kl * tl > (uint32_t)Q_LEV_CAP) tl = (uint32_t)Q_LEV_CAP / kl;
    if (tl == 0) return NSM_ERR_UNSUPPORTED;

    QratioParams p;
    p.L = *left; p.R = *right; p.job = *job;
    p.swap_out = swap_out ? 1u : 0u;
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
        if (p.r_begin < r_lo) p.r_begin = r_lo;
        if (p.r_end > r_hi) p.r_end = r_hi;
        if (p.r_end <= p.r_begin) continue;
        const uint32_t words = w + 1;
        // words actually instantiated: 1, 2, 3, 4, 6, 8
        const uint32_t w_inst = words <= 4 ? words : (words <= 6 ? 6u : 8u);
        uint32_t threads = Q_MAX_THREADS;
        while (threads >= 32 && qratio_layout(p.n_alpha, w_inst, threads, tl, true).total > Q_SMEM_BUDGET)
            threads -= 32;
        if (threads < 32) {
            set_error("alphabet %u x %u words does not fit shared memory", p.n_alpha, w_inst);
            return NSM_ERR_BAD_ARG;
        }
        // no more threads than right items (rounded up to a warp): a small right side, e.g. one
        // term against many synonyms, should not pay for idle mask columns
        const uint32_t n_right = p.r_end - p.r_begin;
        const uint32_t need = ((n_right + 31u) / 32u) * 32u;
        if (threads > need) threads = need;
        p.threads = threads;
        const size_t smem = qratio_layout(p.n_alpha, w_inst, threads, tl, true).total;
        p.n_rblocks = (n_right + threads - 1) / threads;
        const uint64_t n_units = (uint64_t)p.n_lgroups * p.n_rblocks;
        if (n_units > 0xffffffffull) {
            set_error("too many work units (%llu); split the left row block", (unsigned long long)n_units);
            return NSM_ERR_BAD_ARG;
        }
        const uint32_t resident = (uint32_t)sm_count();
        const uint32_t grid = (uint32_t)(n_units < resident ? n_units : resident);
        p.unit_counter = next_unit_counter(stream);
        if (!p.unit_counter) {
            set_error("unit counter: %s", cudaGetErrorString(cudaGetLastError()));
            return NSM_ERR_CUDA;
        }
        if (int rc = dispatch_levels(w_inst, p, smem, grid, stream)) return rc;
    }
    return NSM_OK;
}

// One direction of the product: `left` rows [row_lo, row_hi) x `right` items [r_lo, r_hi), all of
// the right items within the classes <= 8 words (the pattern side of the per-thread kernels).
static int short_pattern_pass(const nsm_strings_t *left, const nsm_strings_t *right, const nsm_job_t *job,
                              uint32_t row_lo, uint32_t row_hi, uint32_t r_lo, uint32_t r_hi, bool levels,
                              bool swap_out, cudaStream_t stream) {
    if (row_hi <= row_lo || r_hi <= r_lo) return NSM_OK;
    nsm_job_t j = *job;
    j.l_row_begin = row_lo; j.l_row_end = row_hi;
    if (swap_out) { j.l_cat = job->r_cat; j.r_cat = job->l_cat; }
    if (!levels) return qratio_flat_launch(left, right, &j, r_lo, r_hi, swap_out, stream);
    // left rows with level strings <= 512 characters are tiled for 512, the (few) others by the
    // side's longest string; an item too large for the tile goes to the generic kernel
    const uint32_t l_short = left->class_end[Q_MAX_WORDS - 1];
    const uint32_t mid = row_lo > l_short ? row_lo : (row_hi < l_short ? row_hi : l_short);
    j.l_row_end = mid;
    const uint32_t cap = 64u * Q_MAX_WORDS;
    int rc = levels_launch(left, right, &j, left->max_len < cap ? left->max_len : cap, r_lo, r_hi, swap_out, stream);
    if (rc) return rc;
    j.l_row_begin = mid; j.l_row_end = row_hi;
    rc = levels_launch(left, right, &j, left->max_len, r_lo, r_hi, swap_out, stream);
    if (rc == NSM_ERR_UNSUPPORTED)
        rc = qratio_long_launch(left, right, &j, mid, row_hi, r_lo, r_hi, swap_out, stream);
    return rc;
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
    if (!left->level_hist || !right->level_hist) {
        set_error("level_hist is missing (pack with gpu/pack.py:pack_strings)");
        return NSM_ERR_BAD_ARG;
    }
    for (int w = 1; w < Q_MAX_WORDS; ++w)
        if (left->class_end[w] < left->class_end[w - 1] || right->class_end[w] < right->class_end[w - 1] ||
            left->class_end[w] > left->n_items || right->class_end[w] > right->n_items) {
            set_error("class_end must be non-decreasing and <= n_items");
            return NSM_ERR_BAD_ARG;
        }
    const bool levels = left->max_levels > 1 || right->max_levels > 1;
    // Items are stored by the length of their longest level string; those from class_end[7] on
    // hold a string of more than 512 characters ("long").  The per-thread kernels keep the RIGHT
    // item's string as the bit-vector pattern (<= 8 words); LCS is symmetric, so
    //   all left rows   x short right items : as they are,
    //   short left rows x long right items  : with the sides swapped (records are swapped back),
    //   long left rows  x long right items  : one warp per pair (qratio_long.cu).
    const uint32_t r_short = right->class_end[Q_MAX_WORDS - 1], l_short = left->class_end[Q_MAX_WORDS - 1];
    const uint32_t lo = job->l_row_begin, hi = job->l_row_end;
    const uint32_t l_short_hi = hi < l_short ? hi : l_short;      // short left rows of the block: [lo, l_short_hi)
    const uint32_t l_long_lo = lo > l_short ? lo : l_short;       // long left rows: [l_long_lo, hi)
    if (int rc = short_pattern_pass(left, right, job, lo, hi, 0u, r_short, levels, false, stream)) return rc;
    if (r_short < right->n_items) {
        if (int rc = short_pattern_pass(right, left, job, r_short, right->n_items, lo, l_short_hi, levels, true,
                                        stream))
            return rc;
        if (l_long_lo < hi)
            if (int rc = qratio_long_launch(left, right, job, l_long_lo, hi, r_short, right->n_items, false, stream))
                return rc;
    }
    return NSM_OK;
}
