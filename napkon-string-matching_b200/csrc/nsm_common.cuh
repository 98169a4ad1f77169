// Shared helpers of the nsm kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nsm.h"

namespace nsm {

void set_error(const char *fmt, ...);
void reset_launch_count();
void count_launch();

#define NSM_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t err__ = (expr);                                                       \
        if (err__ != cudaSuccess) {                                                       \
            nsm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),     \
                           __FILE__, __LINE__);                                           \
            return NSM_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// categories_matching (comparable_data.py:464-476) on per-item category bit masks
__device__ __forceinline__ bool keep_categories(uint32_t mode, uint64_t ml, uint64_t mr) {
    if (mode == NSM_CAT_OFF) return true;
    if (mode == NSM_CAT_LIST_LIST) return (ml & mr) != 0 || (ml == 0 && mr == 0);
    return (ml & mr) != 0;
}

// Threshold compaction: every lane with `keep` appends one record.  One ballot, one
// atomicAdd per warp; records of a warp land next to each other (16 B each -> full sectors).
// Must be called by all 32 lanes of the warp.
__device__ __forceinline__ void emit_pairs(bool keep, uint32_t left, uint32_t right, double score,
                                           nsm_pair_t *__restrict__ out, uint64_t capacity,
                                           unsigned long long *__restrict__ count,
                                           uint32_t *__restrict__ flags) {
    const unsigned m = __ballot_sync(FULL_MASK, keep);
    if (m == 0) return;
    const unsigned leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane_id() == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    if (keep) {
        const unsigned long long pos = base + __popc(m & lanemask_lt());
        if (pos < capacity) {
            // one 16-byte store per record
            double2 rec;
            rec.x = __longlong_as_double((long long)(((unsigned long long)right << 32) | left));
            rec.y = score;
            *reinterpret_cast<double2 *>(out + pos) = rec;
        } else {
            atomicOr(flags, NSM_FLAG_OVERFLOW);
        }
    }
}

// Validates the fields both entry points share and zeroes the counters on the stream.
int prepare_job(const nsm_job_t *job, uint32_t n_left, cudaStream_t stream);

// Largest float that is certainly below every score that could still reach `threshold`
// when scores are accumulated in float64 (see DESIGN.md "filter soundness").
float filter_threshold(double threshold);

int sm_count();

}  // namespace nsm
