// fuzzy_match all-pairs kernel for items with ONE level string each (the flat score function;
// config 3's `Question` strings; MeshProvider.get_matches' term x synonym product), sm_100a.
//
// Replaces the pair loop of ComparableData.gen_comparable (comparable_data.py:223-243) and the
// np.vectorize(fuzzy_match) of terminology/mesh.py:209 for score_func == "fuzzy_match"
// (compare/score_functions.py:20-27): QRatio / 100 = 1 - dist / (m + n) with the Indel distance
// dist = m + n - 2 LCS.
//
// Two things keep most pairs away from the LCS:
//  * an integer keep test.  For a fixed length sum s the float64 map d -> ((1 - d/s) 100)/100 is
//    monotone, so "score >= threshold" is "dist <= dmax[s]" for a table dmax built once per launch
//    WITH that very map (binary search per s): no float64 work per pair, scores are computed for
//    kept pairs only.
//  * a sound lower bound of the distance.  One insertion or deletion changes one character count
//    by one, so dist >= D1 = sum_c |cnt_l(c) - cnt_r(c)|; merging characters into 32 buckets and
//    saturating the counters only lowers D1.  Every level string carries 32 saturating byte
//    counters (bucket = code & 31); D1 is eight VABSDIFF4.ACC instructions, and a pair with
//    D1 > dmax[m + n] is proven below the threshold (this contains the length bound |m - n|).
//
// Work decomposition: as in qratio.cu a thread owns one right item and keeps the pattern-match
// masks of its string in shared memory, transposed [code][word][thread] (conflict-free).  A tile of
// up to 256 left strings is staged in shared memory (padded with a code whose mask row is zero, so
// every string of the tile can be read to the tile's longest length without a predicate).  Phase 1
// tests all tile x thread pairs with the bound and leaves a survivor bit mask per thread.
// Phase 2: every lane walks ITS OWN survivors, two at a time (two independent S chains), reading
// its own text rows; the packer orders both sides by length, so the lanes of a warp hold patterns
// of (nearly) one length, a tile holds texts of (nearly) one length, and survivor counts per lane
// are close to each other.
#include "qratio_common.cuh"

namespace nsm {

constexpr int QF_TILE = 256;            // left strings per tile: survivor counts per lane even out
constexpr int QF_MASK_WORDS = QF_TILE / 64;
constexpr int QF_CHR_CAP = 32 * 1024;   // bytes of left strings staged per batch
constexpr int QF_GROUP = 4;             // tiles per unit (one mask build per unit)

struct QflatParams {
    nsm_strings_t L, R;
    nsm_job_t job;
    uint32_t n_ltiles, n_lgroups, n_rblocks;
    uint32_t threads;   // right items per block
    uint32_t n_rows;    // rows of the mask table: alphabet + the padding code
    uint32_t r_begin, r_end;
    uint32_t table_len; // entries of the dmax table: max length sum + 1
    uint32_t swap_out;  // 1: emit (right, left): the caller swapped the sides (LCS is symmetric)
    uint32_t *unit_counter;  // device counter the CTAs draw their units from; NULL: fixed stride
    double thr_eff;     // threshold on QRatio/100 itself (2 x threshold for compare_terms on K = 1)
};

struct QflatLayout {
    size_t pm, dmax, chr, hist, len, cat, misc, surv, plen, pool, total;
};

__host__ __device__ inline QflatLayout qflat_layout(uint32_t n_rows, uint32_t words, uint32_t threads,
                                                    uint32_t table_len) {
    QflatLayout l;
    size_t o = 0;
    l.pm = o;   o += (size_t)n_rows * words * threads * 8;
    l.chr = o;  o += QF_CHR_CAP;
    l.hist = o; o += QF_TILE * 32;
    l.cat = o;  o += QF_TILE * 8;
    l.len = o;  o += QF_TILE * 4;
    l.misc = o; o += 16;
    l.surv = o; o += (size_t)QF_MASK_WORDS * threads * 8;   // survivor bit masks of the batch, [word][thread]
    l.plen = o; o += (size_t)threads * 4;                    // pattern length of every column
    l.pool = o; o += 16 * 4;                                 // next work word of every bank-pair class
    l.dmax = o; o += ((size_t)table_len * 2 + 15) & ~(size_t)15;
    l.total = (o + 15) & ~(size_t)15;
    return l;
}

// QRatio / 100 of two non-empty processed strings from the Indel distance — the same operation
// sequence as qratio_from_lcs (which it equals for dist = m + n - 2 LCS; empty strings give
// dist == lensum and so 0.0 here too).
__device__ __forceinline__ double qratio_from_dist(uint32_t dist, uint32_t lensum) {
    if (lensum == 0) return 0.0;
    if (lensum <= 64u * Q_MAX_WORDS * 2u) {
        const double norm_dist = div_by_rcp((double)dist, (double)lensum, g_len_rcp.v[lensum]);
        const double norm_sim = __dsub_rn(1.0, norm_dist);
        return div_by_rcp(__dmul_rn(norm_sim, 100.0), 100.0, 0.01);
    }
    const double norm_dist = __ddiv_rn((double)dist, (double)lensum);
    const double norm_sim = __dsub_rn(1.0, norm_dist);
    return __ddiv_rn(__dmul_rn(norm_sim, 100.0), 100.0);
}

// largest distance that still reaches the threshold at this length sum; -1: none
__device__ __forceinline__ int dmax_for(uint32_t lensum, double thr) {
    if (!(qratio_from_dist(0u, lensum) >= thr)) return -1;   // also NaN thresholds
    uint32_t lo = 0, hi = lensum;                            // value(lo) >= thr always holds
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (qratio_from_dist(mid, lensum) >= thr) lo = mid; else hi = mid - 1;
    }
    return (int)lo;
}

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
}

template <int W>
__global__ void __launch_bounds__(Q_MAX_THREADS, 1)
qratio_flat_kernel(const QflatParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31u;
    const QflatLayout lay = qflat_layout(p.n_rows, W, nthr, p.table_len);
    uint64_t *s_pm = reinterpret_cast<uint64_t *>(smem_raw + lay.pm);   // [code][word][thread]
    int16_t *s_dmax = reinterpret_cast<int16_t *>(smem_raw + lay.dmax);
    uint8_t *s_chr = smem_raw + lay.chr;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem_raw + lay.hist);  // [li][8]
    uint32_t *s_len = reinterpret_cast<uint32_t *>(smem_raw + lay.len);    // 0xffffffff: no level
    uint64_t *s_cat = reinterpret_cast<uint64_t *>(smem_raw + lay.cat);
    uint32_t *s_misc = reinterpret_cast<uint32_t *>(smem_raw + lay.misc);  // [0] longest text of the batch
    uint64_t *s_surv = reinterpret_cast<uint64_t *>(smem_raw + lay.surv);
    uint32_t *s_plen = reinterpret_cast<uint32_t *>(smem_raw + lay.plen);
    uint32_t *s_pool = reinterpret_cast<uint32_t *>(smem_raw + lay.pool);

    unsigned long long *count = reinterpret_cast<unsigned long long *>(p.job.out_count);
    nsm_pair_t *out = static_cast<nsm_pair_t *>(p.job.out_pairs);
    const bool flat = p.job.flat != 0;
    const double thr = p.job.threshold;
    const uint32_t pad_code = p.n_rows - 1;
    const uint64_t pad8 = 0x0101010101010101ull * pad_code;
    unsigned long long st_bound = 0, st_cand = 0;
    const uint64_t *pm = s_pm + tid;
    const unsigned char *col = reinterpret_cast<const unsigned char *>(pm);
    const uint32_t word_bytes = nthr * 8u, row_bytes = (uint32_t)W * word_bytes;

    for (uint32_t s = tid; s < p.table_len; s += nthr) {
        const int d = dmax_for(s, p.thr_eff);
        s_dmax[s] = (int16_t)(d > 32767 ? 32767 : d);   // distances above 32767 need strings the tile cannot hold
    }
    // the padding code's mask row stays zero for every pattern
    for (uint32_t x = 0; x < (uint32_t)W; ++x) s_pm[((size_t)pad_code * W + x) * nthr + tid] = 0;

    auto step = [&](uint64_t (&S)[W], uint32_t c, const unsigned char *column) {
        const unsigned char *row = column + c * row_bytes;
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

    // Units are drawn from a device counter (a CTA takes the next one when it is done), and in
    // falling order of cost: both sides are stored by rising length, so the units with the longest
    // patterns and texts come first and the kernel ends on the cheap ones.
    const uint32_t n_units = p.n_lgroups * p.n_rblocks;
    uint32_t unit = blockIdx.x;
    while (unit < n_units) {
        const uint32_t unit_rev = n_units - 1u - unit;
        const uint32_t rb = unit_rev / p.n_lgroups, lgroup = unit_rev - rb * p.n_lgroups;
        const uint32_t r = p.r_begin + rb * nthr + tid;
        const bool r_valid = r < p.r_end;
        uint32_t kr = 0, m = 0;
        uint32_t hr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint64_t rcat = 0;
        __syncthreads();  // nobody still reads masks or the table of the previous unit / the prologue
        if (tid == 0) s_misc[1] = p.unit_counter ? gridDim.x + atomicAdd(p.unit_counter, 1u) : unit + gridDim.x;
        if (r_valid) {
            const uint32_t rg0 = __ldg(p.R.item_level_off + r);
            kr = __ldg(p.R.item_level_off + r + 1) - rg0;
            if (p.job.cat_mode) rcat = __ldg(p.job.r_cat + r);
            if (kr) {
                // my pattern masks (my column only: no barrier needed) and my character counters
                const uint32_t c0 = __ldg(p.R.level_chr_off + rg0);
                m = __ldg(p.R.level_len + rg0);
                for (uint32_t row = 0; row + 1 < p.n_rows; ++row)
#pragma unroll
                    for (int x = 0; x < W; ++x) s_pm[((size_t)row * W + x) * nthr + tid] = 0;
                for (uint32_t j = 0; j < m; j += 8) {
                    const uint64_t w8 = __ldg(reinterpret_cast<const uint64_t *>(p.R.chr + c0 + j));
                    for (uint32_t q = 0; q < 8 && j + q < m; ++q) {
                        const uint32_t c = (uint32_t)(w8 >> (8 * q)) & 0xffu;
                        s_pm[((size_t)c * W + ((j + q) >> 6)) * nthr + tid] |= 1ull << ((j + q) & 63u);
                    }
                }
                const uint4 h0 = __ldg(reinterpret_cast<const uint4 *>(p.R.level_hist) + 2 * (size_t)rg0);
                const uint4 h1 = __ldg(reinterpret_cast<const uint4 *>(p.R.level_hist) + 2 * (size_t)rg0 + 1);
                hr[0] = h0.x; hr[1] = h0.y; hr[2] = h0.z; hr[3] = h0.w;
                hr[4] = h1.x; hr[5] = h1.y; hr[6] = h1.z; hr[7] = h1.w;
            }
        }

        const uint32_t lt_begin = lgroup * QF_GROUP;
        const uint32_t lt_end = min(lt_begin + (uint32_t)QF_GROUP, p.n_ltiles);
        for (uint32_t lt = lt_begin; lt < lt_end; ++lt) {
            const uint32_t t0 = p.job.l_row_begin + lt * QF_TILE;
            const uint32_t tn = min((uint32_t)QF_TILE, p.job.l_row_end - t0);
            // a tile goes through shared memory in one batch, or in several when its strings are long
            for (uint32_t b0 = 0; b0 < tn;) {
                __syncthreads();  // previous batch fully consumed
                if (tid == 0) s_misc[0] = 0;
                if (tid < 16) s_pool[tid] = 0;   // phase 2's work counters (read behind the barriers below)
                __syncthreads();
                // lengths first: they size the batch
                uint32_t my_max = 0;
                for (uint32_t i = tid; i < tn - b0; i += nthr) {
                    const uint32_t g0 = __ldg(p.L.item_level_off + t0 + b0 + i);
                    const uint32_t kl = __ldg(p.L.item_level_off + t0 + b0 + i + 1) - g0;
                    const uint32_t len = kl ? __ldg(p.L.level_len + g0) : 0xffffffffu;
                    s_len[i] = len;
                    if (len != 0xffffffffu) my_max = max(my_max, len);
                }
                my_max = __reduce_max_sync(FULL_MASK, my_max);
                if (lane == 0 && my_max) atomicMax(&s_misc[0], my_max);
                __syncthreads();
                const uint32_t longest = s_misc[0];
                const uint32_t stride8 = (((longest + 7u) >> 3) + 1u) | 1u;  // odd: rows spread over the banks
                const uint32_t stride = stride8 * 8u;
                uint32_t bn = min(tn - b0, (uint32_t)QF_CHR_CAP / stride - 1u);  // one row is the all-padding dummy
                if (bn == 0) { __trap(); }  // the host checked that one string fits
                // batch b0 .. b0+bn: texts (padded), counters, category masks
                for (uint32_t idx = tid; idx < (bn + 1) * stride8; idx += nthr) {
                    const uint32_t li = idx / stride8, w = idx - li * stride8;
                    uint64_t word = pad8;
                    const uint32_t len = li < bn ? s_len[li] : 0u;
                    if (li < bn && len != 0xffffffffu && w * 8u < len) {
                        const uint32_t g0 = __ldg(p.L.item_level_off + t0 + b0 + li);
                        word = __ldg(reinterpret_cast<const uint64_t *>(p.L.chr + __ldg(p.L.level_chr_off + g0)) + w);
                        const uint32_t vb = len - w * 8u;   // valid bytes of this word
                        if (vb < 8u) {
                            const uint64_t keep = (1ull << (8u * vb)) - 1ull;
                            word = (word & keep) | (pad8 & ~keep);
                        }
                    }
                    reinterpret_cast<uint64_t *>(s_chr)[idx] = word;
                }
                for (uint32_t idx = tid; idx < bn * 8u; idx += nthr) {
                    const uint32_t li = idx >> 3;
                    uint32_t v = 0;
                    if (s_len[li] != 0xffffffffu)
                        v = __ldg(p.L.level_hist + 8 * (size_t)__ldg(p.L.item_level_off + t0 + b0 + li) + (idx & 7u));
                    s_hist[idx] = v;
                }
                for (uint32_t li = tid; li < bn; li += nthr)
                    s_cat[li] = p.job.cat_mode ? __ldg(p.job.l_cat + t0 + b0 + li) : 0;
                __syncthreads();
                const uint32_t l0 = t0 + b0;

                // ---- phase 1: the distance bound, all pairs of the batch ----------------------
                uint64_t mask[QF_MASK_WORDS];
#pragma unroll
                for (int x = 0; x < QF_MASK_WORDS; ++x) mask[x] = 0;
                bool special = false;
                uint32_t n_pass = 0;
#pragma unroll
                for (int x = 0; x < QF_MASK_WORDS; ++x) {
                    uint64_t word = 0;
                    const uint32_t li_end = min(bn, 64u * (x + 1));
                    for (uint32_t li = 64u * x; li < li_end; ++li) {
                        const uint32_t n = s_len[li];
                        const uint4 a = reinterpret_cast<const uint4 *>(s_hist)[2 * li];
                        const uint4 b = reinterpret_cast<const uint4 *>(s_hist)[2 * li + 1];
                        uint32_t d1 = sad4(a.x, hr[0], 0u);
                        d1 = sad4(a.y, hr[1], d1); d1 = sad4(a.z, hr[2], d1); d1 = sad4(a.w, hr[3], d1);
                        d1 = sad4(b.x, hr[4], d1); d1 = sad4(b.y, hr[5], d1); d1 = sad4(b.z, hr[6], d1);
                        d1 = sad4(b.w, hr[7], d1);
                        const bool both = n != 0xffffffffu && kr != 0;
                        const bool pass = r_valid && both && keep_categories(p.job.cat_mode, s_cat[li], rcat) &&
                                          (int)d1 <= (int)s_dmax[both ? m + n : 0u];
                        word |= (uint64_t)pass << (li & 63u);
                        special |= (n == 0xffffffffu) || kr == 0;
                    }
                    mask[x] = word;
                    n_pass += __popcll(word);
                }
                st_bound += r_valid ? bn : 0u;
                st_cand += n_pass;
                // items without a level (K == 0): both empty -> compare_terms returns 0; one -> IndexError
                if (__any_sync(FULL_MASK, special)) {
                    for (uint32_t li = 0; li < bn; ++li) {
                        const bool kl0 = s_len[li] == 0xffffffffu;
                        bool ok = r_valid && (kl0 || kr == 0) && keep_categories(p.job.cat_mode, s_cat[li], rcat);
                        if (ok && kl0 != (kr == 0)) { atomicOr(p.job.out_flags, NSM_FLAG_EMPTY_ITEM); ok = false; }
                        emit_pairs(ok && 0.0 >= thr, p.swap_out ? r : l0 + li, p.swap_out ? l0 + li : r, 0.0, out,
                                   p.job.out_capacity, count, p.job.out_flags);
                    }
                }

                // ---- phase 2: the survivors of the whole CTA, pooled per bank-pair class -----------
                // Survivor counts differ up to 7x between the lanes of a warp (a string with rare letters
                // has few histogram neighbours) and the batch ends when its fullest lane does, so the
                // lists are pooled.  A lane may read ANY mask column whose bank pair is its own: columns
                // c with c % 16 == lane % 16 (the 16 lanes of a half-warp then still hit 16 different
                // bank pairs: no conflict).  The survivor words of those columns form the pool of the
                // class; its 2 lanes per warp draw 64-survivor words from it through a shared counter.
                // (the barrier at the top of the batch freed the pool arrays of the previous one)
#pragma unroll
                for (int x = 0; x < QF_MASK_WORDS; ++x) s_surv[(uint32_t)x * nthr + tid] = mask[x];
                s_plen[tid] = m;
                if (__syncthreads_or(n_pass != 0)) {   // (a batch nobody survives, the rule at high thresholds, ends here)
                    const uint32_t cls = lane & 15u;
                    const uint32_t n_words = (nthr >> 4) * (uint32_t)QF_MASK_WORDS;   // words of my class
                    uint64_t cur = 0;        // survivors left in the word I hold
                    uint32_t cur_col = tid, cur_base = 0;
                    bool dry = false;        // the pool of my class is exhausted
                    while (true) {
                        while (cur == 0 && !dry) {
                            const uint32_t idx = atomicAdd(&s_pool[cls], 1u);
                            if (idx >= n_words) { dry = true; break; }
                            const uint32_t x = idx % (uint32_t)QF_MASK_WORDS;
                            cur_col = cls + 16u * (idx / (uint32_t)QF_MASK_WORDS);
                            cur_base = 64u * x;
                            cur = s_surv[x * nthr + cur_col];
                        }
                        if (!__any_sync(FULL_MASK, cur != 0)) break;
                        uint32_t li0 = bn, li1 = bn;   // bn: the all-padding dummy row
                        if (cur) { li0 = cur_base + (uint32_t)__ffsll((long long)cur) - 1u; cur &= cur - 1; }
                        if (cur) { li1 = cur_base + (uint32_t)__ffsll((long long)cur) - 1u; cur &= cur - 1; }
                        const unsigned char *column = reinterpret_cast<const unsigned char *>(s_pm + cur_col);
                        const uint32_t mm = s_plen[cur_col];
                        const uint32_t rr = p.r_begin + rb * nthr + cur_col;
                        const uint32_t n0 = li0 < bn ? s_len[li0] : 0u, n1 = li1 < bn ? s_len[li1] : 0u;
                        const uint32_t trip = __reduce_max_sync(FULL_MASK, max(n0, n1));
                        const uint2 *ta = reinterpret_cast<const uint2 *>(s_chr + li0 * stride);
                        const uint2 *tb = reinterpret_cast<const uint2 *>(s_chr + li1 * stride);
                        uint64_t Sa[W], Sb[W];
#pragma unroll
                        for (int x = 0; x < W; ++x) Sa[x] = Sb[x] = ~0ull;
                        for (uint32_t j = 0; j < trip; j += 8) {
                            const uint2 wa = ta[j >> 3], wb = tb[j >> 3];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                step(Sa, __byte_perm(wa.x, 0u, 0x4440u + q), column);
                                step(Sb, __byte_perm(wb.x, 0u, 0x4440u + q), column);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                step(Sa, __byte_perm(wa.y, 0u, 0x4440u + q), column);
                                step(Sb, __byte_perm(wb.y, 0u, 0x4440u + q), column);
                            }
                        }
                        uint32_t lcs0 = 0, lcs1 = 0;
#pragma unroll
                        for (int x = 0; x < W; ++x) { lcs0 += __popcll(~Sa[x]); lcs1 += __popcll(~Sb[x]); }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const uint32_t li = q ? li1 : li0, n = q ? n1 : n0, lcs = q ? lcs1 : lcs0;
                            // score >= threshold  <=>  dist <= dmax[m + n]  (the table is built with the score's map)
                            const bool keep = li < bn && (int)(mm + n - 2u * lcs) <= (int)s_dmax[li < bn ? mm + n : 0u];
                            double score = 0.0;
                            if (keep) {
                                score = qratio_from_lcs(mm, n, lcs);
                                // compare_terms on K = 1 items: one step, weight 1/2
                                if (!flat) score = __fma_rn(score, 0.5, 0.0);
                            }
                            emit_pairs(keep, p.swap_out ? rr : l0 + li, p.swap_out ? l0 + li : rr, score, out,
                                       p.job.out_capacity, count, p.job.out_flags);
                        }
                    }
                }
                b0 += bn;
            }
        }
        unit = s_misc[1];   // written before the barriers of the tile loop; rewritten behind the next one
    }
    if (p.job.out_stats) {
        for (int o = 16; o; o >>= 1) {
            st_bound += __shfl_xor_sync(FULL_MASK, st_bound, o);
            st_cand += __shfl_xor_sync(FULL_MASK, st_cand, o);
        }
        if (lane == 0) {
            unsigned long long *st = reinterpret_cast<unsigned long long *>(p.job.out_stats);
            if (st_bound) atomicAdd(st + NSM_STAT_BOUND_PAIRS, st_bound);
            if (st_cand) { atomicAdd(st + NSM_STAT_CANDIDATES, st_cand); atomicAdd(st + NSM_STAT_LEVEL_EVALS, st_cand); }
        }
    }
}

template <int W>
static int launch_flat(const QflatParams &p, size_t smem, uint32_t grid, cudaStream_t stream) {
    NSM_CUDA_CHECK(cudaFuncSetAttribute(qratio_flat_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    qratio_flat_kernel<W><<<grid, p.threads, smem, stream>>>(p);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

// All right items of the classes <= Q_MAX_WORDS words x the job's left row block, one launch per
// word class of the right side.  `swap_out`: the caller passed the cohorts swapped.
int qratio_flat_launch(const nsm_strings_t *left, const nsm_strings_t *right, const nsm_job_t *job,
                       uint32_t r_lo, uint32_t r_hi, bool swap_out, cudaStream_t stream) {
    QflatParams p;
    p.L = *left; p.R = *right; p.job = *job;
    p.swap_out = swap_out ? 1u : 0u;
    p.thr_eff = job->flat ? job->threshold : 2.0 * job->threshold;
    p.n_rows = (left->n_alphabet ? left->n_alphabet : 1u) + 1u;
    const uint32_t n_rows_l = job->l_row_end - job->l_row_begin;
    p.n_ltiles = (n_rows_l + QF_TILE - 1) / QF_TILE;
    p.n_lgroups = (p.n_ltiles + QF_GROUP - 1) / QF_GROUP;
    const uint64_t table_len = (uint64_t)left->max_len + 64ull * Q_MAX_WORDS + 1ull;
    const uint32_t row_bytes = ((((left->max_len + 7u) >> 3) + 1u) | 1u) * 8u;
    if (2u * row_bytes > (uint32_t)QF_CHR_CAP || table_len > 32768) {
        set_error("a left level string of %u characters exceeds the staged tile", left->max_len);
        return NSM_ERR_UNSUPPORTED;
    }
    p.table_len = (uint32_t)table_len;
    for (uint32_t w = 0; w < (uint32_t)Q_MAX_WORDS; ++w) {
        p.r_begin = w ? right->class_end[w - 1] : 0u;
        p.r_end = right->class_end[w];
        if (p.r_begin < r_lo) p.r_begin = r_lo;
        if (p.r_end > r_hi) p.r_end = r_hi;
        if (p.r_end <= p.r_begin) continue;
        const uint32_t words = w + 1;
        const uint32_t w_inst = words <= 4 ? words : (words <= 6 ? 6u : 8u);
        uint32_t threads = Q_MAX_THREADS;
        while (threads >= 32 && qflat_layout(p.n_rows, w_inst, threads, p.table_len).total > Q_SMEM_BUDGET)
            threads -= 32;
        if (threads < 32) {
            set_error("alphabet %u x %u words does not fit shared memory", p.n_rows - 1, w_inst);
            return NSM_ERR_UNSUPPORTED;
        }
        const uint32_t n_right = p.r_end - p.r_begin;
        const uint32_t need = ((n_right + 31u) / 32u) * 32u;
        if (threads > need) threads = need;
        p.threads = threads;
        const size_t smem = qflat_layout(p.n_rows, w_inst, threads, p.table_len).total;
        p.n_rblocks = (n_right + threads - 1) / threads;
        const uint64_t n_units = (uint64_t)p.n_lgroups * p.n_rblocks;
        if (n_units > 0xffffffffull) {
            set_error("too many work units (%llu); split the left row block", (unsigned long long)n_units);
            return NSM_ERR_UNSUPPORTED;
        }
        const uint32_t resident = (uint32_t)sm_count();
        const uint32_t grid = (uint32_t)(n_units < resident ? n_units : resident);
        p.unit_counter = next_unit_counter(stream);
        if (!p.unit_counter) {
            set_error("unit counter: %s", cudaGetErrorString(cudaGetLastError()));
            return NSM_ERR_CUDA;
        }
        int rc;
        switch (w_inst) {
            case 1: rc = launch_flat<1>(p, smem, grid, stream); break;
            case 2: rc = launch_flat<2>(p, smem, grid, stream); break;
            case 3: rc = launch_flat<3>(p, smem, grid, stream); break;
            case 4: rc = launch_flat<4>(p, smem, grid, stream); break;
            case 6: rc = launch_flat<6>(p, smem, grid, stream); break;
            default: rc = launch_flat<8>(p, smem, grid, stream); break;
        }
        if (rc) return rc;
    }
    return NSM_OK;
}

}  // namespace nsm
