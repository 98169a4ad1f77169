// Device-side string packing for fuzzy_match: the per-character half of
// napkon_string_matching/gpu/pack.py:pack_strings + text/process.py:default_process, i.e. of what
// rapidfuzz 2.1.x applies inside fuzz.QRatio (/root/reference/napkon_string_matching/compare/
// score_functions.py:27) to the strings join_sorted built (:16-17).
//
// The host maps every DISTINCT code point of a run to the symbol of its processed form (blank
// for non-alphanumerics, else its lower-case form: Python's Unicode tables, a few hundred
// entries) and decides the storage order and offsets of the levels (numpy over levels).  The GPU
// does the work that is per character: the trim, the code of every character, the padding and the
// 32 saturating byte counters of the flat kernel's distance bound.  One warp per level string.
#include "nsm_common.cuh"

namespace nsm {

constexpr int PS_WARPS = 8;

__global__ void __launch_bounds__(PS_WARPS * 32)
pack_strings_measure_kernel(const nsm_raw_strings_t raw, uint32_t *__restrict__ first,
                            uint32_t *__restrict__ len, uint32_t *__restrict__ flags) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t warps = gridDim.x * PS_WARPS;
    bool bad = false;
    for (uint32_t g = blockIdx.x * PS_WARPS + (threadIdx.x >> 5); g < raw.n_levels; g += warps) {
        const uint32_t o = __ldg(raw.level_off + g), n = __ldg(raw.level_off + g + 1) - o;
        uint32_t lo = 0xffffffffu, hi = 0;   // first kept index, one past the last kept index
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t i = base + lane;
            bool keep = false;
            if (i < n) {
                const uint32_t c = __ldg(raw.cps + o + i);
                const uint32_t sym = c < raw.table_len ? __ldg(raw.cp_sym + c) : NSM_STR_SYM_NONE;
                bad |= sym == NSM_STR_SYM_NONE;
                keep = sym != raw.blank_sym;
            }
            const unsigned m = __ballot_sync(FULL_MASK, keep);
            if (m) {
                if (lo == 0xffffffffu) lo = base + (uint32_t)__ffs(m) - 1u;
                hi = base + 32u - (uint32_t)__clz(m);
            }
        }
        if (lane == 0) {
            first[g] = lo == 0xffffffffu ? 0u : lo;
            len[g] = lo == 0xffffffffu ? 0u : hi - lo;
        }
    }
    if (__any_sync(FULL_MASK, bad) && lane == 0) atomicOr(flags, NSM_STR_FLAG_UNMAPPED);
}

__global__ void __launch_bounds__(PS_WARPS * 32)
pack_strings_fill_kernel(const nsm_raw_strings_t raw, const uint32_t *__restrict__ first,
                         const uint32_t *__restrict__ src_level, const uint32_t *__restrict__ level_chr_off,
                         const uint32_t *__restrict__ level_len, const uint8_t *__restrict__ sym_code,
                         uint32_t n_stored, uint8_t *__restrict__ chr, uint32_t *__restrict__ level_hist) {
    __shared__ uint32_t s_cnt[PS_WARPS][32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t warps = gridDim.x * PS_WARPS;
    for (uint32_t s = blockIdx.x * PS_WARPS + warp; s < n_stored; s += warps) {
        const uint32_t g = __ldg(src_level + s), n = __ldg(level_len + s);
        const uint32_t src = __ldg(raw.level_off + g) + __ldg(first + g);
        uint8_t *dst = chr + __ldg(level_chr_off + s);
        s_cnt[warp][lane] = 0;
        __syncwarp();
        const uint32_t padded = (n + 7u) & ~7u;
        for (uint32_t i = lane; i < padded; i += 32) {
            uint32_t code = 0;
            if (i < n) {
                const uint32_t c = __ldg(raw.cps + src + i);
                const uint32_t sym = c < raw.table_len ? __ldg(raw.cp_sym + c) : NSM_STR_SYM_NONE;
                code = sym != NSM_STR_SYM_NONE ? __ldg(sym_code + sym) : 0u;   // (measure flagged it)
                atomicAdd(&s_cnt[warp][code & 31u], 1u);
            }
            dst[i] = (uint8_t)code;
        }
        __syncwarp();
        // bucket b -> byte (b & 3) of word b >> 2, saturating at 255
        if (lane < 8) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) w |= min(s_cnt[warp][4 * lane + b], 255u) << (8 * b);
            level_hist[8 * (size_t)s + lane] = w;
        }
        __syncwarp();
    }
}

static uint32_t grid_for(uint32_t n_warps_wanted) {
    const uint64_t blocks = ((uint64_t)n_warps_wanted + PS_WARPS - 1) / PS_WARPS;
    const uint64_t cap = (uint64_t)sm_count() * 8u;
    return (uint32_t)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace nsm

extern "C" int nsm_pack_strings_measure(const nsm_raw_strings_t *raw, uint32_t *first, uint32_t *len,
                                        uint32_t *flags, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    reset_launch_count();
    if (!raw || !first || !len || !flags || !raw->level_off || !raw->cp_sym || (raw->n_cps && !raw->cps)) {
        set_error("nsm_pack_strings_measure: null argument");
        return NSM_ERR_BAD_ARG;
    }
    NSM_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(uint32_t), stream));
    if (raw->n_levels == 0) return NSM_OK;
    pack_strings_measure_kernel<<<grid_for(raw->n_levels), PS_WARPS * 32, 0, stream>>>(*raw, first, len, flags);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

extern "C" int nsm_pack_strings_fill(const nsm_raw_strings_t *raw, const uint32_t *first,
                                     const uint32_t *src_level, const uint32_t *level_chr_off,
                                     const uint32_t *level_len, const uint8_t *sym_code, uint32_t n_stored,
                                     uint8_t *chr, uint32_t *level_hist, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    reset_launch_count();
    if (!raw || !first || !src_level || !level_chr_off || !level_len || !sym_code || !chr || !level_hist) {
        set_error("nsm_pack_strings_fill: null argument");
        return NSM_ERR_BAD_ARG;
    }
    if (n_stored == 0) return NSM_OK;
    pack_strings_fill_kernel<<<grid_for(n_stored), PS_WARPS * 32, 0, stream>>>(
        *raw, first, src_level, level_chr_off, level_len, sym_code, n_stored, chr, level_hist);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}
