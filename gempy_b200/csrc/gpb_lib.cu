// Library plumbing: error reporting, device query, launch counter.
#include "gpb_common.cuh"
#include <cstdarg>
#include <cstdlib>
#include <mutex>

thread_local char g_gpb_error[512] = "";
std::atomic<long long> g_gpb_launches{0};

int gpb_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_gpb_error, sizeof(g_gpb_error), fmt, ap);
    va_end(ap);
    return code;
}

int gpb_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int gpb_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < GPB_MAX_DEVICES ? dev : GPB_MAX_DEVICES - 1;
}

namespace {
std::mutex g_dev_mutex[GPB_MAX_DEVICES];
std::mutex g_side_init_mutex;
GpbSideStream g_side[GPB_MAX_DEVICES];
int g_side_state[GPB_MAX_DEVICES] = {0};      // 0: not tried, 1: ok, -1: failed
}  // namespace

GpbSideStream* gpb_side_stream() {
    const int dev = gpb_current_device();
    std::lock_guard<std::mutex> g(g_side_init_mutex);
    if (g_side_state[dev] == 0) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        GpbSideStream& sd = g_side[dev];
        const bool ok = cudaStreamCreateWithPriority(&sd.stream, cudaStreamNonBlocking, hi) == cudaSuccess &&
                        cudaEventCreateWithFlags(&sd.ready, cudaEventDisableTiming) == cudaSuccess &&
                        cudaEventCreateWithFlags(&sd.done, cudaEventDisableTiming) == cudaSuccess;
        cudaGetLastError();
        g_side_state[dev] = ok ? 1 : -1;
    }
    return g_side_state[dev] == 1 ? &g_side[dev] : nullptr;
}

// Scratch allocations of the entry points (flags, scan counts, the inverse diagonal blocks of the symmetric solve): a
// per-device stream-ordered pool that KEEPS its memory.  The default pool hands freed memory back to the driver at the next
// synchronisation (release threshold 0), so every call re-mapped a few megabytes; on a shared host that unmap / map pair was
// measured to stall the solve for 50-700 ms in one step out of four (scripts/probes/e2e_breakdown.py).
namespace {
std::mutex g_pool_mutex;
cudaMemPool_t g_pool[GPB_MAX_DEVICES] = {nullptr};
int g_pool_state[GPB_MAX_DEVICES] = {0};      // 0: not tried, 1: ok, -1: failed (fall back to the default pool)
constexpr size_t kPoolMaxBytes = 64u << 20;   // larger requests (a whole system matrix) are not worth keeping
}  // namespace

cudaError_t gpb_malloc_async(void** p, size_t bytes, cudaStream_t s) {
    const int dev = gpb_current_device();
    cudaMemPool_t pool = nullptr;
    if (bytes <= kPoolMaxBytes) {
        std::lock_guard<std::mutex> g(g_pool_mutex);
        if (g_pool_state[dev] == 0 && getenv("GPB_SCRATCH_POOL") && atoi(getenv("GPB_SCRATCH_POOL")) == 0) g_pool_state[dev] = -1;
        if (g_pool_state[dev] == 0) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            unsigned long long keep = ~0ULL;
            const bool ok = cudaMemPoolCreate(&g_pool[dev], &props) == cudaSuccess &&
                            cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
            cudaGetLastError();
            g_pool_state[dev] = ok ? 1 : -1;
        }
        if (g_pool_state[dev] == 1) pool = g_pool[dev];
    }
    return pool ? cudaMallocFromPoolAsync(p, bytes, pool, s) : cudaMallocAsync(p, bytes, s);
}

GpbDeviceLock::GpbDeviceLock() : dev(gpb_current_device()) { g_dev_mutex[dev].lock(); }
GpbDeviceLock::~GpbDeviceLock() { g_dev_mutex[dev].unlock(); }

extern "C" const char* gpb_last_error(void) { return g_gpb_error; }
extern "C" int gpb_version(void) { return 200; }
extern "C" long long gpb_launch_count(void) { return g_gpb_launches.load(); }

extern "C" int gpb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n)
        return gpb_set_error(GPB_E_NODEVICE, "no CUDA device %d (count %d)", device, n);
    int sm = 0, ma = 0, mi = 0;
    GPB_CHECK_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device));
    GPB_CHECK_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, device));
    GPB_CHECK_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, device));
    if (sm_count) *sm_count = sm;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    if (ma != 10) return gpb_set_error(GPB_E_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, ma, mi);
    return GPB_OK;
}

// ---- FP64 FMA pipe microbenchmark (roofline denominator; MEASURED_PEAKS.json has no FP64 entry) ----------
__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;      // never true: keeps the chain alive
}

extern "C" int gpb_bench_dfma(int iters, double* tflops_host, void* stream) {
    GPB_REQUIRE(iters > 0 && tflops_host, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    double* d = nullptr;
    GPB_CHECK_CUDA(cudaMalloc((void**)&d, sizeof(double)));
    const int blocks = gpb_sm_count() * 8;
    cudaEvent_t e0, e1;
    GPB_CHECK_CUDA(cudaEventCreate(&e0));
    GPB_CHECK_CUDA(cudaEventCreate(&e1));
    dfma_chain_kernel<<<blocks, 256, 0, s>>>(d, iters / 10 + 1, 0.999999, 1e-9);   // warm-up
    GPB_LAUNCH_CHECK();
    GPB_CHECK_CUDA(cudaEventRecord(e0, s));
    dfma_chain_kernel<<<blocks, 256, 0, s>>>(d, iters, 0.999999, 1e-9);
    GPB_LAUNCH_CHECK();
    GPB_CHECK_CUDA(cudaEventRecord(e1, s));
    GPB_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GPB_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
    *tflops_host = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return GPB_OK;
}

// ---- FP64 tensor-core (DMMA m8n8k4) microbenchmark: is mma.sync f64 faster than the DFMA pipe on this part? ----
__global__ void __launch_bounds__(256) dmma_chain_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = c[i][0] + 0.5; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

extern "C" int gpb_bench_dmma(int iters, double* tflops_host, void* stream) {
    GPB_REQUIRE(iters > 0 && tflops_host, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    double* d = nullptr;
    GPB_CHECK_CUDA(cudaMalloc((void**)&d, sizeof(double)));
    const int blocks = gpb_sm_count() * 8;
    cudaEvent_t e0, e1;
    GPB_CHECK_CUDA(cudaEventCreate(&e0));
    GPB_CHECK_CUDA(cudaEventCreate(&e1));
    dmma_chain_kernel<<<blocks, 256, 0, s>>>(d, iters / 10 + 1, 1e-3, 1e-3);
    GPB_LAUNCH_CHECK();
    GPB_CHECK_CUDA(cudaEventRecord(e0, s));
    dmma_chain_kernel<<<blocks, 256, 0, s>>>(d, iters, 1e-3, 1e-3);
    GPB_LAUNCH_CHECK();
    GPB_CHECK_CUDA(cudaEventRecord(e1, s));
    GPB_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GPB_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    // one m8n8k4 = 8*8*4 FMA = 512 flop per warp instruction
    const double flops = 512.0 * 8.0 * (double)iters * 8.0 * (double)blocks;
    *tflops_host = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return GPB_OK;
}

// ---- do DFMA and DMMA overlap?  Every warp interleaves `ratio` DFMA per DMMA; reports both rates. ----------
template <int RATIO>
__global__ void __launch_bounds__(256) mixed_chain_kernel(double* out, int iters, double a, double b) {
    double c[4][2];
    double x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = c[i][0] + 0.5; }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int r = 0; r < RATIO; ++r) x[(i * RATIO + r) & 7] = fma(x[(i * RATIO + r) & 7], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

extern "C" int gpb_bench_mixed(int iters, int ratio, double* dfma_tflops_host, double* dmma_tflops_host, void* stream) {
    GPB_REQUIRE(iters > 0 && dfma_tflops_host && dmma_tflops_host, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    double* d = nullptr;
    GPB_CHECK_CUDA(cudaMalloc((void**)&d, sizeof(double)));
    const int blocks = gpb_sm_count() * 8;
    cudaEvent_t e0, e1;
    GPB_CHECK_CUDA(cudaEventCreate(&e0));
    GPB_CHECK_CUDA(cudaEventCreate(&e1));
    auto run = [&](int it) {
        switch (ratio) {
            case 4: mixed_chain_kernel<4><<<blocks, 256, 0, s>>>(d, it, 0.999999, 1e-9); break;
            case 8: mixed_chain_kernel<8><<<blocks, 256, 0, s>>>(d, it, 0.999999, 1e-9); break;
            case 16: mixed_chain_kernel<16><<<blocks, 256, 0, s>>>(d, it, 0.999999, 1e-9); break;
            default: mixed_chain_kernel<32><<<blocks, 256, 0, s>>>(d, it, 0.999999, 1e-9); break;
        }
    };
    run(iters / 10 + 1);
    GPB_LAUNCH_CHECK();
    GPB_CHECK_CUDA(cudaEventRecord(e0, s));
    run(iters);
    GPB_LAUNCH_CHECK();
    GPB_CHECK_CUDA(cudaEventRecord(e1, s));
    GPB_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GPB_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const int r = (ratio == 4 || ratio == 8 || ratio == 16) ? ratio : 32;
    const double warps = 8.0 * blocks;
    *dmma_tflops_host = 512.0 * 4.0 * iters * warps / (ms * 1e-3) / 1e12;
    *dfma_tflops_host = 64.0 * 4.0 * r * (double)iters * warps / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return GPB_OK;
}
