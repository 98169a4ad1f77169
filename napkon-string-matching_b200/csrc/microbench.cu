// Integer-pipe micro-benchmarks: the measured denominators of the integer roofline
// (SURVEY.md §8d asks for a measured IADD3/LOP3 and POPC stream, not a spec number).
#include "nsm_common.cuh"

namespace nsm {

constexpr int MB_CHAINS = 8;

template <int KIND>
__global__ void microbench_kernel(uint32_t iters, uint32_t *__restrict__ sink) {
    uint32_t x[MB_CHAINS];
    const uint32_t seed = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int c = 0; c < MB_CHAINS; ++c) x[c] = seed * 2654435761u + c * 40503u + 1u;
    const uint32_t k1 = seed | 0x10001u, k2 = ~seed;
    uint64_t y[MB_CHAINS / 2];
#pragma unroll
    for (int c = 0; c < MB_CHAINS / 2; ++c) y[c] = ((uint64_t)x[2 * c] << 32) | x[2 * c + 1];
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int c = 0; c < MB_CHAINS; ++c) {
                if (KIND == 0) {  // one LOP3 (three-input logic)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                } else if (KIND == 1) {  // one IADD3
                    asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }"
                                 : "+r"(x[c]) : "r"(k1), "r"(k2));
                } else if (KIND == 2) {  // one POPC (+ one IADD to keep a dependency)
                    x[c] = __popc(x[c]) + k1;
                }
            }
            if (KIND == 3) {  // Hyyro LCS step on a 64-bit word: u = S & M; S = (S + u) | (S - u)
#pragma unroll
                for (int c = 0; c < MB_CHAINS / 2; ++c) {
                    const uint64_t m = ((uint64_t)k1 << 32 | k2) ^ (it + rep);
                    const uint64_t u = y[c] & m;
                    y[c] = (y[c] + u) | (y[c] - u);
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < MB_CHAINS; ++c) acc ^= x[c];
#pragma unroll
    for (int c = 0; c < MB_CHAINS / 2; ++c) acc ^= (uint32_t)y[c] ^ (uint32_t)(y[c] >> 32);
    if (acc == 0x12345678u) sink[0] = acc;  // practically never; keeps the chains alive
}

}  // namespace nsm

extern "C" int nsm_microbench(int kind, uint32_t blocks, uint32_t threads, uint32_t iters,
                              uint32_t *sink, uint64_t *ops_per_thread, void *stream_) {
    using namespace nsm;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!sink || blocks == 0 || threads == 0 || threads > 1024) {
        set_error("bad microbench launch");
        return NSM_ERR_BAD_ARG;
    }
    uint64_t per_iter = 0;
    reset_launch_count();
    count_launch();
    switch (kind) {
        case 0: microbench_kernel<0><<<blocks, threads, 0, stream>>>(iters, sink); per_iter = 4 * MB_CHAINS; break;
        case 1: microbench_kernel<1><<<blocks, threads, 0, stream>>>(iters, sink); per_iter = 4 * MB_CHAINS; break;  // one IADD3 each
        case 2: microbench_kernel<2><<<blocks, threads, 0, stream>>>(iters, sink); per_iter = 4 * MB_CHAINS; break;
        // one LCS step = AND, ADD, SUB, OR on 64 bits = 8 int32 ops
        case 3: microbench_kernel<3><<<blocks, threads, 0, stream>>>(iters, sink); per_iter = 4 * (MB_CHAINS / 2) * 8; break;
        default: set_error("unknown microbench kind %d", kind); return NSM_ERR_BAD_ARG;
    }
    NSM_CUDA_CHECK(cudaGetLastError());
    if (ops_per_thread) *ops_per_thread = per_iter * iters;
    return NSM_OK;
}
