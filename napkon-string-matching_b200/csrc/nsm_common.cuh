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

// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP) completing on an mbarrier ----------------
// One thread arms the barrier with the byte count and issues the copies; everybody waits on the
// barrier's phase.  Sizes and both addresses must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                              uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// orders earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Waits for the phase with the given parity.  Bounded: a barrier that never completes (a bug)
// traps after ~seconds instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) __trap();
    }
}

// Validates the fields both entry points share and zeroes the counters on the stream.
int prepare_job(const nsm_job_t *job, uint32_t n_left, cudaStream_t stream);

// Largest float that is certainly below every score that could still reach `threshold`
// when scores are accumulated in float64 (see DESIGN.md "filter soundness").
float filter_threshold(double threshold);

int sm_count();

// A zeroed device counter for the launch about to be queued on `stream` (see api.cu); nullptr on a
// CUDA error.
uint32_t *next_unit_counter(cudaStream_t stream);

}  // namespace nsm
