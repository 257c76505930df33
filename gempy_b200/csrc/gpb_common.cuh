// Shared helpers of the gempy_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../include/gempy_b200.h"

// Numerical constants pinned by the reference's approved vectors (see oracle/gempy_oracle.py header).
#define GPB_REG_EPS  1e-5     // h_u h_v / (r^2 + 1e-5)
#define GPB_DIST_EPS 1e-10    // r = sqrt(|h|^2 + 1e-10)

#include <atomic>
extern thread_local char g_gpb_error[512];
extern std::atomic<long long> g_gpb_launches;

int gpb_set_error(int code, const char* fmt, ...);

#define GPB_CHECK_CUDA(expr)                                                                       \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return gpb_set_error(GPB_E_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,  \
                                 cudaGetErrorString(_e));                                          \
    } while (0)

#define GPB_LAUNCH_CHECK()                                                                         \
    do {                                                                                           \
        ++g_gpb_launches;                                                                          \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return gpb_set_error(GPB_E_CUDA, "kernel launch failed at %s:%d: %s", __FILE__,        \
                                 __LINE__, cudaGetErrorString(_e));                                \
    } while (0)

#define GPB_REQUIRE(cond, msg)                                                                     \
    do {                                                                                           \
        if (!(cond)) return gpb_set_error(GPB_E_INVALID, "%s (%s:%d)", msg, __FILE__, __LINE__);   \
    } while (0)

static inline long long gpb_round_up(long long v, long long m) { return (v + m - 1) / m * m; }

int gpb_sm_count();
// Index of the current CUDA device clamped to [0, GPB_MAX_DEVICES): per-device one-time state (function attributes,
// side streams) is kept in arrays indexed by it, so one process may drive several devices.
#define GPB_MAX_DEVICES 64
int gpb_current_device();

// High-priority side stream + two events of the current device (panel look-ahead of the LU / Cholesky
// factorisations).  Created on first use, one per device, never destroyed; nullptr if the creation failed.
struct GpbSideStream {
    cudaStream_t stream;
    cudaEvent_t ready, done;
};
GpbSideStream* gpb_side_stream();
// cudaMallocAsync from a per-device pool that keeps its memory (release with cudaFreeAsync); see gpb_lib.cu
cudaError_t gpb_malloc_async(void** p, size_t bytes, cudaStream_t s);
// The side stream and its events are shared by every caller on a device: the host-side ENQUEUE of a factorisation holds
// this lock (event record / wait pairs are resolved at enqueue time, so serialising the enqueues is sufficient).
struct GpbDeviceLock {
    GpbDeviceLock();
    ~GpbDeviceLock();
    int dev;
};

// ---- internal evaluation entry (gpb_eval.cu), used by the C ABI wrappers and by the level executor (gpb_model.cu) ----
struct GpbEvalCall {
    const gpb_stack* st = nullptr;
    const double* src = nullptr;          // packed evaluation table
    int regular = 0;                      // 1: points i0 .. i0 + m of `grid`; 0: explicit points xyz [3][ld_xyz]
    gpb_regular_grid grid{};
    long long i0 = 0;
    const double* xyz = nullptr;
    long long ld_xyz = 0;
    long long m = 0;
    const long long* m_dev = nullptr;     // explicit points only: actual count on the device (m is then the upper bound)
    int octets = 0;                       // explicit points only: points 8g .. 8g+7 are sibling octets (centre +- quarter cell)
    const double* fault_vals = nullptr;   // row f: fault_vals + (fault_ids ? fault_ids[f] : f) * ld_fault, minus fault_min[row]
    long long ld_fault = 0;
    const int* fault_ids = nullptr;
    const double* fault_min = nullptr;
    double* Z = nullptr;
    double* gx = nullptr;
    double* gy = nullptr;
    double* gz = nullptr;
    double* block = nullptr;              // fused activator output (optional)
    const double* act_iso = nullptr;
    const double* act_ids = nullptr;
    int act_n = 0;
    double act_slope = 0.0;
    double* block_min = nullptr;          // running minimum of block (optional)
};
int gpb_eval_call(const GpbEvalCall& c, cudaStream_t stream);

// ---- fast FP64 primitives -------------------------------------------------------------------------
// MUFU seeds (2^-22) + one third-order correction: error ~ e^3 ~ 1e-20 relative before rounding,
// with no divergent slow path (arguments are strictly positive, normal numbers on this path).
// MUFU seeds.  ncu shows the XU pipe (MUFU.RSQ64H) 70 % busy next to a 65 % busy FP64 pipe in the field-only evaluation
// kernels (profiles/r2_eval_octet_kernel_ncu_full_cfg4.txt), so a variant that narrows the double to a float with three
// integer instructions, uses the 32-bit MUFU and widens the result back (same 2^-22 accuracy) was tried: it measured SLOWER
// (512^3 benchmark 789 -> 854 ms per step, config-5 octree levels 1.15 -> 1.48 s) -- the five extra integer instructions per
// pair cost more issue slots and latency than the shorter XU occupancy saves.  Kept behind GPB_MUFU32 for the record.
// Mixing the two (32-bit seeds for the reciprocal of the orientation gradient term only, or for 2 / 4 of the 8 points a
// thread owns) was slower in proportion to the share: 787 -> 803 / 804 / 831 ms.  Field-only kernels (10.5 FP64 instructions per
// surface-point pair) behave the same: 256^3 z-run 57.36 ms with 64-bit seeds, 58.33 / 58.52 / 59.45 ms with 1 / 2 / 3 of the 8
// points on 32-bit seeds -- that kernel already issues 97 % of the FP64 instructions per second of the DFMA-chain probe, so the
// XU pipe is not what binds it, whatever its `xu_realtime` figure suggests.
__device__ __forceinline__ double gpb_rsqrt_seed64(double u) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
    return y;
}
__device__ __forceinline__ double gpb_rcp_seed64(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    return y;
}
__device__ __forceinline__ float gpb_narrow(double u) {
    const unsigned hi = (unsigned)__double2hiint(u) - 0x38000000u;          // exponent bias 1023 -> 127
    return __uint_as_float(__funnelshift_l((unsigned)__double2loint(u), hi, 3));
}
__device__ __forceinline__ double gpb_widen(float y) {
    const unsigned b = __float_as_uint(y);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
__device__ __forceinline__ double gpb_rsqrt_seed32(double u) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(gpb_narrow(u)));
    return gpb_widen(y);
}
__device__ __forceinline__ double gpb_rcp_seed32(double d) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(gpb_narrow(d)));
    return gpb_widen(y);
}
#ifndef GPB_MUFU32
__device__ __forceinline__ double gpb_rsqrt_seed(double u) { return gpb_rsqrt_seed64(u); }
__device__ __forceinline__ double gpb_rcp_seed(double d) { return gpb_rcp_seed64(d); }
#else
__device__ __forceinline__ double gpb_rsqrt_seed(double u) { return gpb_rsqrt_seed32(u); }
__device__ __forceinline__ double gpb_rcp_seed(double d) { return gpb_rcp_seed32(d); }
#endif
// sqrt(u), u > 0
__device__ __forceinline__ double gpb_fast_sqrt(double u) {
    const double y0 = gpb_rsqrt_seed(u);
    const double g0 = u * y0;
    const double e = fma(-g0, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(g0 * e, p, g0);
}
// Second-order variants for the evaluation kernel (one FP64 instruction less each): relative error ~ 3/8 e^2 = 2e-14
// (sqrt) and e^2 = 6e-14 (reciprocal) for the 2^-22 MUFU seeds -- five orders inside the 1e-9 field tolerance, and the
// pair terms they enter are summed with alternating signs.  The covariance assembly (1e-12 matrix tolerance) keeps the
// third-order forms.
__device__ __forceinline__ double gpb_fast_sqrt2(double u) {
    const double y0 = gpb_rsqrt_seed(u);
    const double g0 = u * y0;
#ifndef GPB_SQRT2_INT_HALF
    const double e = fma(-g0, y0, 1.0);
    return g0 * fma(0.5, e, 1.0);
#else
    // g0 + (u - g0^2) * y0/2 with the halving as an exponent decrement on the integer pipe: three FP64 instructions
    // instead of four, same error term.  Measured SLOWER on the 512^3 benchmark (779.7 -> 787.4 ms per step): an integer
    // instruction in the pair loop costs as much issue bandwidth as the FP64 instruction it replaces.
    const double h = __hiloint2double(__double2hiint(y0) - 0x00100000, __double2loint(y0));
    return fma(fma(-g0, g0, u), h, g0);
#endif
}
__device__ __forceinline__ double gpb_fast_rcp2(double d) {
    const double y0 = gpb_rcp_seed(d);
    const double e = fma(-d, y0, 1.0);
    return fma(y0, e, y0);
}
// 1/d, d > 0
__device__ __forceinline__ double gpb_fast_rcp(double d) {
    const double y0 = gpb_rcp_seed(d);
    const double e = fma(-d, y0, 1.0);
    const double p = fma(e, e, e);
    return fma(y0, p, y0);
}

// exp(x) for x <= 0 in the pair loops of the exponential and Matern kernels: 2^k e^r with k = rint(x log2 e) taken from the
// low word of x log2 e + 1.5 * 2^52, |r| <= ln2 / 2 by a two-term Cody-Waite reduction, degree-12 Taylor polynomial
// (relative error 5e-16 in exact arithmetic, < 2e-15 measured against libm over [-700, 0]), the scaling as an exponent
// addition.  No table, no branch: libdevice's exp() costs 19 FP64 and ~15 integer / predicate / branch instructions per call
// (SASS of the Matern pair loop), this one 18 FP64 and 3 integer.  Arguments below -700 return exp(-700) ~ 1e-304.
__device__ __forceinline__ double gpb_fast_exp_neg(double x) {
    x = fmax(x, -700.0);
    const double kMagic = 6755399441055744.0;
    const double kf = fma(x, 1.4426950408889634074, kMagic);
    const int k = __double2loint(kf);
    const double kd = kf - kMagic;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 2.08767569878680989792e-09;            // 1/12!
    p = fma(p, r, 2.50521083854417187751e-08);        // 1/11!
    p = fma(p, r, 2.75573192239858906526e-07);        // 1/10!
    p = fma(p, r, 2.75573192239858906526e-06);        // 1/9!
    p = fma(p, r, 2.48015873015873015873e-05);        // 1/8!
    p = fma(p, r, 1.98412698412698412698e-04);        // 1/7!
    p = fma(p, r, 1.38888888888888888889e-03);        // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);        // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);        // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);        // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// ---- mbarrier / bulk-copy (TMA) helpers -------------------------------------------------------------
__device__ __forceinline__ uint32_t gpb_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void gpb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gpb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gpb_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void gpb_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void gpb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gpb_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void gpb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(gpb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void gpb_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     gpb_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(gpb_smem_u32(bar))
                 : "memory");
}
