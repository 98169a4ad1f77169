// Entry points that are not kernels: version, error text, job validation, device facts.
#include <algorithm>
#include <atomic>
#include <vector>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "nsm_common.cuh"

namespace nsm {

static thread_local char g_error[512] = "";
static thread_local int g_launches = 0;

void reset_launch_count() { g_launches = 0; }
void count_launch() { ++g_launches; }
int launch_count() { return g_launches; }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

float filter_threshold(double threshold) {
    if (isnan(threshold)) return nanf("");
    if (!(threshold > 0.0)) return -INFINITY;
    // the float64 accumulation can exceed the real-valued sum by a few ulp; 1e-9 relative slack
    // covers that by many orders of magnitude, then step one fp32 value down so the conversion's
    // own rounding cannot lift the bound
    const float f = (float)(threshold * (1.0 - 1e-9));
    return nextafterf(f, -INFINITY);
}

int prepare_job(const nsm_job_t *job, uint32_t n_left, cudaStream_t stream) {
    reset_launch_count();
    if (job->l_row_begin > job->l_row_end || job->l_row_end > n_left) {
        set_error("row block [%u, %u) outside the %u left items", job->l_row_begin, job->l_row_end,
                  n_left);
        return NSM_ERR_BAD_ARG;
    }
    if (!job->out_count || !job->out_flags || (job->out_capacity && !job->out_pairs)) {
        set_error("out_count / out_flags / out_pairs must be device pointers");
        return NSM_ERR_BAD_ARG;
    }
    if (job->cat_mode > NSM_CAT_MEMBER || (job->cat_mode && (!job->l_cat || !job->r_cat))) {
        set_error("bad category mode or missing category masks");
        return NSM_ERR_BAD_ARG;
    }
    if (job->out_mode > NSM_OUT_CODED) {
        set_error("unknown out_mode %u", job->out_mode);
        return NSM_ERR_BAD_ARG;
    }
    if (job->out_mode == NSM_OUT_CODED &&
        (!job->out_dict || !job->out_exc_count || (job->out_exc_capacity && !job->out_exc) ||
         (reinterpret_cast<uintptr_t>(job->out_exc) & 15u))) {
        set_error("NSM_OUT_CODED needs out_dict, out_exc_count and a 16-byte aligned out_exc");
        return NSM_ERR_BAD_ARG;
    }
    if (reinterpret_cast<uintptr_t>(job->out_pairs) & 15u) {
        set_error("out_pairs must be 16-byte aligned");
        return NSM_ERR_BAD_ARG;
    }
    NSM_CUDA_CHECK(cudaMemsetAsync(job->out_count, 0, sizeof(uint64_t), stream));
    NSM_CUDA_CHECK(cudaMemsetAsync(job->out_flags, 0, sizeof(uint32_t), stream));
    if (job->out_stats)
        NSM_CUDA_CHECK(cudaMemsetAsync(job->out_stats, 0, NSM_N_STATS * sizeof(uint64_t), stream));
    if (job->out_mode == NSM_OUT_CODED)
        NSM_CUDA_CHECK(cudaMemsetAsync(job->out_exc_count, 0, sizeof(uint64_t), stream));
    return NSM_OK;
}

// Unit counters: a small per-device pool, one entry per launch in flight (round robin), zeroed on
// the launch's stream right before it.  Persistent kernels draw their work units from them.
constexpr int N_UNIT_COUNTERS = 4096;   // far more than the launches a process keeps in flight
__device__ uint32_t g_unit_counters[N_UNIT_COUNTERS];

uint32_t *next_unit_counter(cudaStream_t stream) {
    static std::atomic<uint32_t> ticket{0};
    uint32_t *base = nullptr;
    if (cudaGetSymbolAddress(reinterpret_cast<void **>(&base), g_unit_counters) != cudaSuccess) return nullptr;
    uint32_t *ctr = base + ticket.fetch_add(1) % N_UNIT_COUNTERS;
    if (cudaMemsetAsync(ctr, 0, sizeof(uint32_t), stream) != cudaSuccess) return nullptr;
    return ctr;
}

}  // namespace nsm

extern "C" int nsm_version(void) { return NSM_VERSION; }

extern "C" const char *nsm_last_error(void) { return nsm::g_error; }

extern "C" int nsm_last_launch_count(void) { return nsm::g_launches; }

extern "C" int nsm_dict_reset(uint64_t *dict, void *stream) {
    if (!dict) { nsm::set_error("null dictionary"); return NSM_ERR_BAD_ARG; }
    NSM_CUDA_CHECK(cudaMemsetAsync(dict, 0xff, NSM_DICT_SLOTS * sizeof(uint64_t),
                                   static_cast<cudaStream_t>(stream)));
    return NSM_OK;
}

namespace nsm {
__global__ void publish_kernel(const unsigned long long *__restrict__ src, volatile unsigned long long *dst,
                               uint32_t words) {
    if (threadIdx.x < words) dst[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}
}  // namespace nsm

extern "C" int nsm_publish(const void *dev_src, void *host_dst, uint32_t bytes, void *stream) {
    using namespace nsm;
    reset_launch_count();
    if (!dev_src || !host_dst || bytes == 0 || bytes > 256 || (bytes & 7u) ||
        (reinterpret_cast<uintptr_t>(dev_src) & 7u) || (reinterpret_cast<uintptr_t>(host_dst) & 7u)) {
        set_error("nsm_publish: 8-byte aligned pointers and 8..256 bytes in multiples of 8");
        return NSM_ERR_BAD_ARG;
    }
    publish_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const unsigned long long *>(dev_src), static_cast<volatile unsigned long long *>(host_dst),
        bytes / 8);
    count_launch();
    NSM_CUDA_CHECK(cudaGetLastError());
    return NSM_OK;
}

// ---- host side of the compact record formats ---------------------------------------------------
extern "C" uint64_t nsm_decode_packets(const nsm_packet_t *packets, uint64_t n_packets,
                                       const uint32_t *left_perm, const uint32_t *right_perm,
                                       nsm_pair_t *out) {
    uint64_t n = 0;
    for (uint64_t i = 0; i < n_packets; ++i) {
        const nsm_packet_t &pk = packets[i];
        const uint32_t cnt = pk.count < NSM_PACKET_RECORDS ? pk.count : NSM_PACKET_RECORDS;
        for (uint32_t k = 0; k < cnt; ++k, ++n) {
            const uint32_t l = pk.left0 + (pk.local[k] >> 7), r = pk.right0 + (pk.local[k] & 127u);
            out[n].left = left_perm ? left_perm[l] : l;
            out[n].right = right_perm ? right_perm[r] : r;
            out[n].score = pk.score[k];
        }
    }
    return n;
}

extern "C" uint64_t nsm_decode_cpackets(const nsm_cpacket_t *packets, uint64_t n_packets, const uint64_t *dict,
                                        const uint32_t *left_perm, const uint32_t *right_perm,
                                        nsm_pair_t *out) {
    uint64_t n = 0;
    for (uint64_t i = 0; i < n_packets; ++i) {
        const nsm_cpacket_t &pk = packets[i];
        const uint32_t cnt = pk.count < NSM_CPACKET_RECORDS ? pk.count : NSM_CPACKET_RECORDS;
        for (uint32_t k = 0; k < cnt; ++k, ++n) {
            const uint32_t rec = pk.rec[k];
            const uint32_t l = pk.left0 + ((rec & 0xffffu) >> 7), r = pk.right0 + (rec & 127u);
            out[n].left = left_perm ? left_perm[l] : l;
            out[n].right = right_perm ? right_perm[r] : r;
            uint64_t bits = dict[rec >> 16];
            memcpy(&out[n].score, &bits, sizeof(double));
        }
    }
    return n;
}

extern "C" int nsm_sort_pairs(const nsm_pair_t *records, uint64_t n, uint32_t n_left, nsm_pair_t *sorted) {
    if ((n && (!records || !sorted)) || (records && sorted && records == sorted)) {
        nsm::set_error("nsm_sort_pairs: null or overlapping buffers");
        return NSM_ERR_BAD_ARG;
    }
    std::vector<uint64_t> start((size_t)n_left + 1, 0);
    for (uint64_t i = 0; i < n; ++i) {
        if (records[i].left >= n_left) {
            nsm::set_error("nsm_sort_pairs: left index %u >= n_left %u", records[i].left, n_left);
            return NSM_ERR_BAD_ARG;
        }
        ++start[(size_t)records[i].left + 1];
    }
    for (size_t l = 0; l < n_left; ++l) start[l + 1] += start[l];
    std::vector<uint64_t> fill(start.begin(), start.end() - 1);
    for (uint64_t i = 0; i < n; ++i) sorted[fill[records[i].left]++] = records[i];
    for (size_t l = 0; l < n_left; ++l)
        if (start[l + 1] - start[l] > 1)
            std::sort(sorted + start[l], sorted + start[l + 1],
                      [](const nsm_pair_t &a, const nsm_pair_t &b) { return a.right < b.right; });
    return NSM_OK;
}
