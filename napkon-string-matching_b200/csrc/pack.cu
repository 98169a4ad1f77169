// Device-side token packing (sm_100a): builds the arrays of nsm_sets_t from dictionary codes.
//
// The host tokenises items and maps token strings to integer codes (Q3: exact string identity);
// everything after that - the level sets of ComparableData.gen_comp_value
// (/root/reference/napkon_string_matching/types/comparable_data.py:283-285: level j = tokens of the
// item's last j+1 parts), sort + unique per level, frequency-ranked ids, head / tail signatures,
// the compare_terms slot schedule (:248-265) - is integer work per item and runs here, one warp
// per item.  The output is bit-identical to napkon_string_matching/gpu/pack.py:finish_sets.
//
// One warp, one item:
//   1. every id of the item becomes a 64-bit key  id << 32 | level << 16  (parts: level = index
//      of its part counted from the back, i.e. the first level that contains the part);
//   2. the keys are sorted in shared memory (bitonic, warp-synchronous);
//   3. one ordered sweep drops duplicates (parts: one row per id, its entry level = the smallest
//      level; explicit levels: one row per (id, level)) and stamps each row with the entry level
//      of its id (low 16 bits), compacting in place with ballot + popc;
//   4. level j is a filter of that array (parts: entry <= j; levels: level == j) and comes out
//      sorted by id; its bitsets are OR-reduced over the warp and lanes t = 1..n_slots scatter
//      the summary into the slot rows t with min(t, K-1) == j.
// Two passes (measure, fill) with an exclusive scan of the per-item row counts in between.
#include "nsm_common.cuh"

namespace nsm {

constexpr int PK_WARPS = 4;
constexpr uint32_t PK_CAP = NSM_PACK_MAX_ITEM_IDS;
constexpr uint32_t PK_HEAD = 64;  // pack.py HEAD_IDS
constexpr uint32_t PK_MULT1 = 0x9E3779B1u, PK_MULT2 = 0x85EBCA77u;  // pack.py SIG_HASH_MULT, SIG2_HASH_MULT
constexpr int SC_THREADS = 1024;

struct PackParams {
    nsm_raw_sets_t raw;
    nsm_sets_t out;                // fill pass: arrays to write
    const uint32_t *item_tok_off;  // fill pass: first token row of every item
    uint32_t *item_count;          // measure pass: token rows of every item
    uint32_t *flags;
};

__device__ __forceinline__ uint64_t warp_or64(uint64_t v) {
    const uint32_t lo = __reduce_or_sync(FULL_MASK, (uint32_t)v);
    const uint32_t hi = __reduce_or_sync(FULL_MASK, (uint32_t)(v >> 32));
    return ((uint64_t)hi << 32) | lo;
}

// ascending bitonic sort of n2 (a power of two >= 32) keys in shared memory by one warp
__device__ __forceinline__ void warp_bitonic_sort(uint64_t *keys, uint32_t n2, unsigned lane) {
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < n2; i += 32) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const uint64_t a = keys[i], b = keys[x];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[x] = a; }
                }
            }
            __syncwarp();
        }
    }
}

template <bool FILL>
__global__ void __launch_bounds__(PK_WARPS * 32) pack_sets_kernel(const PackParams p) {
    __shared__ uint64_t s_keys[PK_WARPS][PK_CAP];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t *keys = s_keys[warp];
    const bool parts = p.raw.mode == NSM_RAW_SUFFIX_PARTS;
    const bool exact = p.out.exact_bits != 0;
    const uint32_t n_slots = p.out.n_slots, stride = p.out.slot_stride;
    uint32_t *o_tok = const_cast<uint32_t *>(p.out.tok);
    uint8_t *o_entry = const_cast<uint8_t *>(p.out.tok_entry);
    uint32_t bad = 0;

    if (FILL && blockIdx.x == 0 && threadIdx.x == 0)
        const_cast<uint32_t *>(p.out.level_tok_off)[p.raw.n_groups] = __ldg(p.item_tok_off + p.raw.n_items);

    for (uint32_t item = blockIdx.x * PK_WARPS + warp; item < p.raw.n_items; item += gridDim.x * PK_WARPS) {
        const uint32_t g0 = __ldg(p.raw.item_grp_off + item), g1 = __ldg(p.raw.item_grp_off + item + 1);
        const uint32_t K = g1 - g0;
        const uint32_t e0 = __ldg(p.raw.grp_id_off + g0), e1 = __ldg(p.raw.grp_id_off + g1);
        const uint32_t n = e1 - e0;
        if (n > PK_CAP || K > 0xffffu) {
            bad |= NSM_PACK_FLAG_TOO_LARGE;
            if (!FILL && lane == 0) p.item_count[item] = 0;
            continue;
        }
        // ---- 1. keys -----------------------------------------------------------------------
        for (uint32_t e = lane; e < n; e += 32) {
            const uint32_t pos = e0 + e;
            uint32_t id = __ldg(p.raw.ids + pos);
            if (id >= p.raw.n_vocab) { bad |= NSM_PACK_FLAG_BAD_ID; id = 0; }
            if (p.raw.rank) id = __ldg(p.raw.rank + id);
            // the group that holds row pos: grp_id_off[lo] <= pos < grp_id_off[lo + 1]
            uint32_t lo = g0, hi = g1;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(p.raw.grp_id_off + mid) <= pos) lo = mid; else hi = mid;
            }
            const uint32_t j = lo - g0;
            const uint32_t lvl = parts ? K - 1 - j : j;
            keys[e] = ((uint64_t)id << 32) | ((uint64_t)lvl << 16);
        }
        uint32_t n2 = 32;
        while (n2 < n) n2 <<= 1;
        for (uint32_t e = n + lane; e < n2; e += 32) keys[e] = ~0ull;  // sorts behind every real key
        __syncwarp();
        // ---- 2. sort -----------------------------------------------------------------------
        warp_bitonic_sort(keys, n2, lane);
        // ---- 3. unique rows, each stamped with the entry level of its id -----------------------
        uint32_t m = 0;           // rows kept so far (warp-uniform)
        uint64_t prev = ~0ull;    // key before the current chunk
        uint32_t carry_entry = 0;
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t pidx = base + lane;
            const bool valid = pidx < n;
            const uint64_t key = valid ? keys[pidx] : ~0ull;
            uint64_t before = __shfl_up_sync(FULL_MASK, key, 1);
            if (lane == 0) before = prev;
            prev = __shfl_sync(FULL_MASK, key, 31);
            const uint32_t lvl = (uint32_t)(key >> 16) & 0xffffu;
            const bool new_id = valid && (pidx == 0 || (before >> 32) != (key >> 32));
            const bool new_key = valid && (pidx == 0 || (before >> 16) != (key >> 16));
            const unsigned starts = __ballot_sync(FULL_MASK, new_id);
            const unsigned mine = starts & (lanemask_lt() | (1u << lane));
            uint32_t entry = __shfl_sync(FULL_MASK, lvl, mine ? 31 - __clz(mine) : 0);
            if (!mine) entry = carry_entry;  // my id's run began in an earlier chunk
            carry_entry = __shfl_sync(FULL_MASK, entry, 31);
            const bool keep = parts ? new_id : new_key;
            const unsigned km = __ballot_sync(FULL_MASK, keep);
            __syncwarp();  // the whole chunk is in registers: rows may now be overwritten
            if (keep) keys[m + __popc(km & lanemask_lt())] = (key & ~0xffffull) | entry;
            m += __popc(km);
            __syncwarp();
        }
        // explicit levels: nested iff every row below the deepest level is followed by the row
        // (same id, level + 1), i.e. an id stays once it has entered
        if (!parts) {
            for (uint32_t base = 0; base < m; base += 32) {
                const uint32_t pidx = base + lane;
                if (pidx < m) {
                    const uint64_t key = keys[pidx];
                    const uint32_t lvl = (uint32_t)(key >> 16) & 0xffffu;
                    if (lvl + 1 < K) {
                        const uint64_t nxt = pidx + 1 < m ? keys[pidx + 1] : ~0ull;
                        if ((nxt >> 32) != (key >> 32) || ((uint32_t)(nxt >> 16) & 0xffffu) != lvl + 1)
                            bad |= NSM_PACK_FLAG_NOT_NESTED;
                    }
                }
            }
        }
        // ---- 4. levels ---------------------------------------------------------------------
        const uint32_t tok_base = FILL ? __ldg(p.item_tok_off + item) : 0u;
        uint32_t total = 0;
        uint64_t any_h = 0, any_t = 0;
        for (uint32_t j = 0; j < K; ++j) {
            uint32_t size = 0;
            uint64_t head = 0, tail = 0, tail2 = 0;
            for (uint32_t base = 0; base < m; base += 32) {
                const uint32_t pidx = base + lane;
                const uint64_t key = pidx < m ? keys[pidx] : 0ull;
                const uint32_t id = (uint32_t)(key >> 32), lvl = (uint32_t)(key >> 16) & 0xffffu,
                               entry = (uint32_t)key & 0xffffu;
                const bool sel = pidx < m && (parts ? entry <= j : lvl == j);
                const unsigned sm = __ballot_sync(FULL_MASK, sel);
                if (FILL && sel) {
                    const uint32_t at = tok_base + total + size + __popc(sm & lanemask_lt());
                    o_tok[at] = id;
                    o_entry[at] = (uint8_t)min(entry, 255u);
                    if (id < PK_HEAD) {
                        head |= 1ull << id;
                    } else {
                        tail |= 1ull << (exact ? id - PK_HEAD : (id * PK_MULT1) >> 26);
                        tail2 |= 1ull << (exact ? id - PK_HEAD : (id * PK_MULT2) >> 26);
                    }
                }
                size += __popc(sm);
            }
            if (FILL) {
                head = warp_or64(head); tail = warp_or64(tail); tail2 = warp_or64(tail2);
                const uint32_t folded = size - __popcll(head) - __popcll(tail);
                const uint32_t info = size | (min(folded, 255u) << 16);
                const uint32_t g = g0 + j;
                if (lane == 0) {
                    const_cast<uint32_t *>(p.out.level_tok_off)[g] = tok_base + total;
                    const_cast<uint64_t *>(p.out.level_head)[g] = head;
                    const_cast<uint64_t *>(p.out.level_tail)[g] = tail;
                    const_cast<uint64_t *>(p.out.level_tail2)[g] = tail2;
                    const_cast<uint32_t *>(p.out.level_info)[g] = info;
                }
                // compare_terms' schedule: step t reads level min(t, K-1)
                const uint32_t t = lane + 1;
                if (t <= n_slots && min(t, K - 1) == j) {
                    const size_t at = (size_t)(t - 1) * stride + item;
                    reinterpret_cast<ulonglong2 *>(const_cast<uint64_t *>(p.out.slot_ht))[at] =
                        make_ulonglong2(head, tail);
                    const_cast<uint32_t *>(p.out.slot_info)[at] = info;
                }
                if (j >= 1 || K == 1) { any_h |= head; any_t |= tail; }
            }
            total += size;
        }
        if (lane == 0) {
            if (FILL) {
                reinterpret_cast<ulonglong2 *>(const_cast<uint64_t *>(p.out.item_any))[item] =
                    make_ulonglong2(any_h, any_t);
                const_cast<uint32_t *>(p.out.item_k)[item] = K;
            } else {
                p.item_count[item] = total;
            }
        }
        __syncwarp();  // keys are reused by the next item
    }
    if (bad) atomicOr(p.flags, bad);
}

// ---- exclusive scan of the per-item row counts (three small kernels, 1024 items per tile) ----
__device__ __forceinline__ uint64_t block_inclusive_scan(uint64_t v, uint64_t *warp_tot, uint64_t &block_total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t y = __shfl_up_sync(FULL_MASK, x, d);
        if ((int)lane >= d) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const uint64_t w = warp_tot[lane];
        uint64_t xs = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(FULL_MASK, xs, d);
            if ((int)lane >= d) xs += y;
        }
        warp_tot[lane] = xs - w;  // exclusive prefix of the warp totals
        if (lane == 31) warp_tot[32] = xs;
    }
    __syncthreads();
    x += warp_tot[warp];
    block_total = warp_tot[32];
    __syncthreads();  // warp_tot may be reused
    return x;
}

__global__ void __launch_bounds__(SC_THREADS) scan_tiles_kernel(uint32_t *data, uint64_t *tile_sum, uint32_t n) {
    __shared__ uint64_t warp_tot[33];
    const uint32_t i = blockIdx.x * SC_THREADS + threadIdx.x;
    const uint64_t v = i < n ? data[i] : 0ull;
    uint64_t total;
    const uint64_t x = block_inclusive_scan(v, warp_tot, total);
    if (i < n) data[i] = (uint32_t)(x - v);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SC_THREADS) scan_tile_sums_kernel(uint64_t *tile_sum, uint32_t n_tiles,
                                                                     uint64_t *totals) {
    __shared__ uint64_t warp_tot[33];
    uint64_t carry = 0;
    for (uint32_t base = 0; base < n_tiles; base += SC_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n_tiles ? tile_sum[i] : 0ull;
        uint64_t total;
        const uint64_t x = block_inclusive_scan(v, warp_tot, total);
        if (i < n_tiles) tile_sum[i] = carry + x - v;
        carry += total;
    }
    if (threadIdx.x == 0) totals[0] = carry;
}

__global__ void __launch_bounds__(SC_THREADS) scan_add_kernel(uint32_t *data, const uint64_t *tile_sum,
                                                               const uint64_t *totals, uint32_t n) {
    const uint32_t i = blockIdx.x * SC_THREADS + threadIdx.x;
    if (i < n) data[i] += (uint32_t)tile_sum[blockIdx.x];
    if (i == n) data[n] = (uint32_t)totals[0];
}

__global__ void count_ids_kernel(const uint32_t *__restrict__ ids, uint32_t n_ids, uint32_t n_vocab,
                                 uint32_t *__restrict__ counts) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ids; i += gridDim.x * blockDim.x) {
        const uint32_t id = __ldg(ids + i);
        if (id < n_vocab) atomicAdd(counts + id, 1u);
    }
}

static int check_raw(const nsm_raw_sets_t *raw) {
    if (!raw || !raw->item_grp_off || !raw->grp_id_off || (raw->n_ids && !raw->ids)) {
        set_error("null argument");
        return NSM_ERR_BAD_ARG;
    }
    if (raw->mode > NSM_RAW_LEVELS) {
        set_error("unknown raw mode %u", raw->mode);
        return NSM_ERR_BAD_ARG;
    }
    return NSM_OK;
}

static uint32_t pack_grid(uint32_t n_items) {
    const uint32_t need = (n_items + PK_WARPS - 1) / PK_WARPS;
    const uint32_t resident = (uint32_t)sm_count() * 6u;  // 32 KB of keys per CTA
    return need < resident ? (need ? need : 1u) : resident;
}

}  // namespace nsm

extern "C" int nsm_pack_count_ids(const nsm_raw_sets_t *raw, uint32_t *counts, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    reset_launch_count();
    if (int rc = check_raw(raw)) return rc;
    if (!counts) { set_error("null argument"); return NSM_ERR_BAD_ARG; }
    if (raw->n_ids == 0) return NSM_OK;
    const uint32_t blocks = (uint32_t)sm_count() * 8u;
    count_ids_kernel<<<blocks, 256, 0, stream>>>(raw->ids, raw->n_ids, raw->n_vocab, counts);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

extern "C" uint64_t nsm_pack_scratch_bytes(uint32_t n_items) {
    return ((uint64_t)(n_items + 1u + nsm::SC_THREADS - 1) / nsm::SC_THREADS + 1u) * sizeof(uint64_t);
}

extern "C" int nsm_pack_sets_measure(const nsm_raw_sets_t *raw, uint32_t *item_tok_off, uint64_t *totals,
                                     void *scratch, uint64_t scratch_bytes, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    reset_launch_count();
    if (int rc = check_raw(raw)) return rc;
    if (!item_tok_off || !totals || !scratch || scratch_bytes < nsm_pack_scratch_bytes(raw->n_items)) {
        set_error("item_tok_off / totals / scratch missing or scratch too small");
        return NSM_ERR_BAD_ARG;
    }
    NSM_CUDA_CHECK(cudaMemsetAsync(totals, 0, 2 * sizeof(uint64_t), stream));
    NSM_CUDA_CHECK(cudaMemsetAsync(item_tok_off, 0, ((size_t)raw->n_items + 1) * sizeof(uint32_t), stream));
    if (raw->n_items == 0) return NSM_OK;
    PackParams p{};
    p.raw = *raw;
    p.item_count = item_tok_off;
    p.flags = reinterpret_cast<uint32_t *>(totals + 1);
    pack_sets_kernel<false><<<pack_grid(raw->n_items), PK_WARPS * 32, 0, stream>>>(p);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    // exclusive scan in place; entry n_items receives the total
    const uint32_t n = raw->n_items;
    const uint32_t n_tiles = (n + 1u + SC_THREADS - 1) / SC_THREADS;  // covers index n as well
    uint64_t *tile_sum = static_cast<uint64_t *>(scratch);
    scan_tiles_kernel<<<n_tiles, SC_THREADS, 0, stream>>>(item_tok_off, tile_sum, n);
    scan_tile_sums_kernel<<<1, SC_THREADS, 0, stream>>>(tile_sum, n_tiles, totals);
    scan_add_kernel<<<n_tiles, SC_THREADS, 0, stream>>>(item_tok_off, tile_sum, totals, n);
    count_launch(); count_launch(); count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

extern "C" int nsm_pack_sets_fill(const nsm_raw_sets_t *raw, const uint32_t *item_tok_off,
                                  const nsm_sets_t *out, uint32_t *flags, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    reset_launch_count();
    if (int rc = check_raw(raw)) return rc;
    if (!item_tok_off || !out || !flags) { set_error("null argument"); return NSM_ERR_BAD_ARG; }
    if (out->n_items != raw->n_items || out->n_levels != raw->n_groups) {
        set_error("out describes %u items / %u levels, raw has %u / %u", out->n_items, out->n_levels,
                  raw->n_items, raw->n_groups);
        return NSM_ERR_BAD_ARG;
    }
    if (out->n_slots < 1 || out->n_slots > 32 || out->slot_stride < out->n_items || (out->slot_stride & 127u)) {
        set_error("n_slots must be in 1..32 and slot_stride a multiple of 128 >= n_items");
        return NSM_ERR_BAD_ARG;
    }
    if (!out->item_level_off || !out->level_tok_off || !out->tok || !out->tok_entry || !out->level_head ||
        !out->level_tail || !out->level_tail2 || !out->level_info || !out->item_any || !out->item_k ||
        !out->slot_ht || !out->slot_info) {
        set_error("every array of out must be allocated");
        return NSM_ERR_BAD_ARG;
    }
    NSM_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(uint32_t), stream));
    NSM_CUDA_CHECK(cudaMemsetAsync(const_cast<uint64_t *>(out->slot_ht), 0,
                                   (size_t)out->n_slots * out->slot_stride * 16, stream));
    NSM_CUDA_CHECK(cudaMemsetAsync(const_cast<uint32_t *>(out->slot_info), 0,
                                   (size_t)out->n_slots * out->slot_stride * 4, stream));
    NSM_CUDA_CHECK(cudaMemcpyAsync(const_cast<uint32_t *>(out->item_level_off), raw->item_grp_off,
                                   ((size_t)raw->n_items + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                   stream));
    PackParams p{};
    p.raw = *raw;
    p.out = *out;
    p.item_tok_off = item_tok_off;
    p.flags = flags;
    pack_sets_kernel<true><<<pack_grid(raw->n_items), PK_WARPS * 32, 0, stream>>>(p);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}
