// (2b) Symmetric solve of the saddle-point co-kriging system  [K U; U^T 0] [w; mu] = [b_K; b_u].
//
// Engine stage replaced: "solver" with kernel_solver = 1 (the reference calls a general dense solve,
// numpy.linalg.solve, and ignores the structure) -- SURVEY.md 8a2 row (2); call site gempy/API/compute_API.py:68-73.
// The covariance block K (orientation-gradient, cross and surface-point increment covariances + nuggets) is
// symmetric positive definite; only the universal-drift and fault-drift rows make the system indefinite.  So instead
// of a pivoted LU of the whole matrix (2/3 n^3 flop, one latency-bound pivot search per column) this path runs
//
//   1. a blocked right-looking Cholesky K = L L^T on the LOWER triangle, no pivoting (1/3 n^3 flop):
//        potrf_diag_kernel   64 x 64 diagonal block in shared memory (one CTA)
//        trsm_panel_kernel   L21 = A21 L11^-T, one matrix row per thread held in registers
//        syrk_kernel         A22 -= L21 L21^T on the FP64 tensor cores (mma.sync m8n8k4 DMMA), lower tiles only
//      in outer blocks of 128 columns (the big update runs with K = 128: 16 flop per byte of C traffic), the next
//      outer block being factored on a high-priority side stream while the main stream finishes the update;
//   2. the drift rows U^T and the right-hand sides b^T ride along as extra ROWS below K: the same trsm / syrk turn
//      them into (L^-1 U)^T and (L^-1 b)^T (forward substitution for free) and leave -(U^T K^-1 U) in the corner;
//   3. schur_kernel: the small Schur system for mu (Cholesky of U^T K^-1 U), then z = y - Y_u mu;
//   4. trsv_lt_kernel: L^T w = z, one CTA per 64-row block, pipelined through release/acquire flags.
//
// If K is not numerically positive definite (a pivot <= 0) `info` reports the column and the host mirror falls back
// to the pivoted LU (gpb_lu_solve).
#include "gpb_common.cuh"
#include <cstdlib>

namespace {

constexpr int kCB = 64;            // panel width
constexpr int kCOuter = 128;       // outer block = K of the big trailing update
constexpr int kMaxNu = 64;         // drift + fault rows

// ---- right-hand sides as extra rows: A[n + r, j] = b[j, r] -------------------------------------------------------------
__global__ void set_rhs_rows_kernel(int n, double* __restrict__ A, int lda, const double* __restrict__ b, int nrhs, int ldb) {
    const long long total = (long long)n * nrhs;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / n), j = (int)(e - (long long)r * n);
        A[(long long)j * lda + n + r] = b[(long long)r * ldb + j];
    }
}

__global__ void zero_int_kernel(int* p) { if (p) *p = 0; }

// ---- Cholesky of a small lower-triangular block held in shared memory (column-major, leading dimension ld) ------------
// One barrier per column: the rank-1 update reads the UNSCALED column j (both factors multiplied by 1/sqrt(d) on the
// fly), the column is scaled afterwards; nobody reads column j again.
// A pivot counts as non-positive when d <= kPivotTol * (the original diagonal entry, diag0): a covariance block whose
// pivot lost 13 digits to cancellation is numerically singular and is better served by the pivoted LU.
constexpr double kPivotTol = 1e-13;
__device__ __forceinline__ void chol_smem(double* S, int ld, int m, int* info, int info_base, const double* diag0) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int j = 0; j < m; ++j) {
        const double d = S[j * ld + j];
        const bool bad = !(d > (diag0 ? kPivotTol * fabs(diag0[j]) : 0.0));
        if (bad && tid == 0 && info) atomicCAS(info, 0, info_base + j + 1);
        const double dd = bad ? 1.0 : d;
        const double inv = rsqrt(dd);
        for (int c = j + 1 + warp; c < m; c += nw) {
            const double lc = S[j * ld + c] * inv;
            for (int i = c + lane; i < m; i += 32) {
                const double li = S[j * ld + i] * inv;
                S[c * ld + i] = fma(-li, lc, S[c * ld + i]);
            }
        }
        __syncthreads();
        for (int i = j + 1 + tid; i < m; i += nt) S[j * ld + i] *= inv;
        if (tid == 0) S[j * ld + j] = dd * inv;
    }
    __syncthreads();
}

constexpr int kPLd = kCB + 1;
__global__ void save_diag_kernel(int n, const double* __restrict__ A, int lda, double* __restrict__ diag0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) diag0[i] = A[(long long)i * lda + i];
}

// 1/sqrt(u), u > 0 normal: MUFU seed (2^-22) + one third-order step (error ~1e-20 before rounding)
__device__ __forceinline__ double rsqrt_fast(double u) {
    const double y0 = gpb_rsqrt_seed(u);
    const double g = u * y0;
    const double e = fma(-g, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y0 * e, p, y0);
}

// ---- diagonal block: L11 = chol(A11) and Wt = L11^-T, one CTA ------------------------------------------------------------
// The 64 x 64 block lives in shared memory with 64 IDENTITY rows appended below it: the column eliminations that turn the
// matrix rows into L11 turn the identity rows into L11^-T (rows riding along = forward substitution).  Blocked by 16
// columns: (i) warp 0 factors the 16 x 16 diagonal block in registers, one matrix row per lane (lanes 0-15) and the
// block's 16 identity rows in lanes 16-31, pivots and multipliers exchanged by shuffles -- no block barrier inside the
// 16 sequential columns; (ii) every remaining row times the 16 x 16 inverse that (i) just produced; (iii) rank-16 update
// of the block's remaining columns.  The panel below is then a GEMM with Wt (trsm_dmma_kernel).
constexpr int kPB = 16;
constexpr int kLdS = 129;                      // 64 matrix rows + 64 identity rows + 1
constexpr size_t kPotrfSmem = (size_t)(kCB * kLdS + kCB) * sizeof(double);

__global__ void __launch_bounds__(256) potrf64_kernel(int k0, int jb, double* __restrict__ A, int lda, int* __restrict__ info,
                                                      const double* __restrict__ diag0, double* __restrict__ Wt, double pivot_tol) {
    extern __shared__ double S[];              // S[c * kLdS + slot]
    double* tol = S + kCB * kLdS;              // pivot thresholds of the block's columns
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (info != nullptr && *reinterpret_cast<volatile int*>(info) != 0) return;      // an earlier pivot failed: the caller falls back to the LU
    for (int e = tid; e < kCB * kLdS; e += 256) S[e] = 0.0;
    if (tid < kCB) tol[tid] = (tid < jb && diag0) ? pivot_tol * fabs(diag0[k0 + tid]) : 0.0;
    __syncthreads();
    {
        double tmp[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int e = tid + 256 * t, c = e >> 6, i = e & 63;
            tmp[t] = (c < jb && i < jb && i >= c) ? A[(long long)(k0 + c) * lda + k0 + i] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int e = tid + 256 * t, c = e >> 6, i = e & 63;
            S[c * kLdS + i] = tmp[t];
        }
        if (tid < jb) S[tid * kLdS + 64 + tid] = 1.0;
    }
    __syncthreads();
    const int nbk = (jb + kPB - 1) / kPB;
    for (int b = 0; b < nbk; ++b) {
        const int o = b * kPB;
        const int w = min(kPB, jb - o);
        // (i) 16 x 16 diagonal block + its 16 identity rows, in registers of warp 0
        if (warp == 0) {
            const int ridx = (lane < 16) ? o + lane : 64 + o + (lane - 16);
            double v[kPB];
#pragma unroll
            for (int c = 0; c < kPB; ++c) v[c] = S[(o + c) * kLdS + ridx];
#pragma unroll
            for (int j = 0; j < kPB; ++j) {
                if (j < w) {
                    const double d = __shfl_sync(0xffffffffu, v[j], j);
                    const bool bad = !(d > tol[o + j]);
                    if (bad && lane == 0 && info) atomicCAS(info, 0, k0 + o + j + 1);
                    const double inv = rsqrt_fast(bad ? 1.0 : d);
                    const double l = v[j] * inv;
                    v[j] = l;
#pragma unroll
                    for (int c = j + 1; c < kPB; ++c) {
                        const double lc = __shfl_sync(0xffffffffu, l, c);
                        v[c] = fma(-l, lc, v[c]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kPB; ++c)
                if (lane >= 16 || c <= lane) S[(o + c) * kLdS + ridx] = v[c];
        }
        __syncthreads();
        // (ii) rows below the diagonal block and the identity rows of the earlier blocks: X = rows * L_bb^-T
        const int nm = max(0, jb - o - kPB);
        if (tid < nm + o) {
            const int i = (tid < nm) ? o + kPB + tid : 64 + (tid - nm);
            double a[kPB], x[kPB];
#pragma unroll
            for (int k = 0; k < kPB; ++k) a[k] = S[(o + k) * kLdS + i];
#pragma unroll
            for (int n = 0; n < kPB; ++n) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k <= n; ++k) acc = fma(a[k], S[(o + n) * kLdS + 64 + o + k], acc);
                x[n] = acc;
            }
#pragma unroll
            for (int n = 0; n < kPB; ++n) S[(o + n) * kLdS + i] = x[n];
        }
        __syncthreads();
        // (iii) rank-16 update of the block's remaining columns (two threads per row slot)
        const int c_lo = o + kPB;
        if (c_lo < jb) {
            const int slot = tid >> 1, half = tid & 1;
            const bool valid = (slot < 64) ? (slot >= c_lo && slot < jb) : (slot - 64 < c_lo);
            if (valid) {
                double xi[kPB];
#pragma unroll
                for (int k = 0; k < kPB; ++k) xi[k] = S[(o + k) * kLdS + slot];
                const int c_hi = (slot < 64) ? min(jb, slot + 1) : jb;
                for (int c = c_lo + half; c < c_hi; c += 2) {
                    double acc = S[c * kLdS + slot];
#pragma unroll
                    for (int k = 0; k < kPB; ++k) acc = fma(-xi[k], S[(o + k) * kLdS + c], acc);
                    S[c * kLdS + slot] = acc;
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int e = tid + 256 * t, c = e >> 6, i = e & 63;
        if (c < jb && i < jb && i >= c) A[(long long)(k0 + c) * lda + k0 + i] = S[c * kLdS + i];
    }
    for (int e = tid; e < kCB * kCB; e += 256) {       // Wt[k][n] = (L11^-T)[k][n], zero below the diagonal / beyond jb
        const int k = e >> 6, n = e & 63;
        Wt[e] = (k <= n && n < jb) ? S[n * kLdS + 64 + k] : 0.0;
    }
}

// ---- panel below the diagonal block: L21 = A21 * Wt on the FP64 tensor cores, 64 rows per CTA, in place ------------------
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b);
__device__ __forceinline__ void cpa8(double* dst, const double* src, bool valid);
__device__ __forceinline__ void cpa16(double* dst, const double* src, int bytes);
constexpr int kTLd = kCB + 4;
constexpr size_t kTrsmSmem = (size_t)2 * kCB * kTLd * sizeof(double);

template <bool kAligned16>
__global__ void __launch_bounds__(128) trsm_dmma_kernel(int M, int k0, int jb, double* __restrict__ A, int lda,
                                                        const double* __restrict__ Wt) {
    extern __shared__ __align__(16) double smem_trsm[];
    double* As = smem_trsm;                    // As[k * kTLd + m]
    double* Bs = smem_trsm + kCB * kTLd;       // Bs[k * kTLd + n]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = k0 + jb + (int)blockIdx.x * kCB;
    const double* Ap = A + (long long)k0 * lda;
    if constexpr (kAligned16) {
#pragma unroll
        for (int t = 0; t < kCB * kCB / 2 / 128; ++t) {
            const int e = tid + 128 * t;
            const int k = e >> 5, m = 2 * (e & 31);
            const int gm = row0 + m;
            const int bytes = (k < jb) ? (gm + 1 < M ? 16 : (gm < M ? 8 : 0)) : 0;
            cpa16(As + k * kTLd + m, bytes ? Ap + (long long)k * lda + gm : Ap, bytes);
        }
    } else {
#pragma unroll
        for (int t = 0; t < kCB * kCB / 128; ++t) {
            const int e = tid + 128 * t;
            const int k = e >> 6, m = e & 63;
            const bool ok = (k < jb) && (row0 + m < M);
            cpa8(As + k * kTLd + m, ok ? Ap + (long long)k * lda + row0 + m : Ap, ok);
        }
    }
#pragma unroll
    for (int t = 0; t < kCB * kCB / 2 / 128; ++t) {
        const int e = tid + 128 * t;
        const int k = e >> 5, n = 2 * (e & 31);
        cpa16(Bs + k * kTLd + n, Wt + k * kCB + n, 16);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    const int r = lane >> 2, q = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int kend = wn + 32;                  // Wt is upper triangular: column n needs k <= n only
    for (int ks = 0; ks < kend; ks += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[(ks + q) * kTLd + wm + 8 * i + r];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[(ks + q) * kTLd + wn + 8 * j + r];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gm = row0 + wm + 8 * i + r;
            const int gn = wn + 8 * j + 2 * q;
            if (gm < M && gn < jb) A[(long long)(k0 + gn) * lda + gm] = acc[i][j][0];
            if (gm < M && gn + 1 < jb) A[(long long)(k0 + gn + 1) * lda + gm] = acc[i][j][1];
        }
}

// ---- trailing update on the FP64 tensor cores: C[i, j] -= sum_k X[i, k] X[j, k], i >= j ----------------------------------
// X = the factored panel (columns [src, src + K) of A, rows below it); C = columns [col_begin, col_end) of A, rows
// [col_begin, M).  128 x 64 CTA tiles, tiles entirely above the diagonal exit at once; 8 warps of 32 x 32; K streamed in
// chunks of 32 through a double-buffered cp.async ring; both operands come from the same panel (the B operand is the
// panel's rows at the tile's COLUMN indices).  Epilogue: kRed = false loads C into the accumulators up front (as -C) and
// stores -acc; kRed = true sends -acc as fire-and-forget FP64 reductions to the L2.
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa8(double* dst, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa16(double* dst, const double* src, int bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

constexpr int kSM = 128, kSN = 64, kSK = 32, kSStages = 2, kSThreads = 256;
constexpr int kSLdA = kSM + 4, kSLdB = kSN + 4;           // ld = 4 mod 16 doubles: conflict-free fragment loads per half-warp
constexpr int kSStageDoubles = kSK * (kSLdA + kSLdB);
constexpr size_t kSyrkSmem = (size_t)kSStages * kSStageDoubles * sizeof(double);

template <bool kAligned16, bool kRed>
__global__ void __launch_bounds__(kSThreads, 2)
syrk_kernel(int M, int K, const double* __restrict__ X, double* __restrict__ A, int lda, int col_begin, int col_end) {
    const int r0 = col_begin + (int)blockIdx.x * kSM;       // first row of the tile
    const int c0 = col_begin + (int)blockIdx.y * kSN;       // first column of the tile
    if (r0 + kSM - 1 < c0) return;                          // entirely above the diagonal
    extern __shared__ __align__(16) double smem_syrk[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = (K + kSK - 1) / kSK;

    auto stage_load = [&](int slot, int chunk) {
        double* As = smem_syrk + (size_t)slot * kSStageDoubles;
        double* Bs = As + kSK * kSLdA;
        const int kc = chunk * kSK;
        if constexpr (kAligned16) {
#pragma unroll
            for (int t = 0; t < kSK * kSM / 2 / kSThreads; ++t) {
                const int e = tid + kSThreads * t;
                const int k = e / (kSM / 2), m = 2 * (e % (kSM / 2));
                const int gm = r0 + m;
                const int bytes = (kc + k < K) ? (gm + 1 < M ? 16 : (gm < M ? 8 : 0)) : 0;
                cpa16(As + k * kSLdA + m, bytes ? X + (long long)(kc + k) * lda + gm : X, bytes);
            }
#pragma unroll
            for (int t = 0; t < kSK * kSN / 2 / kSThreads; ++t) {
                const int e = tid + kSThreads * t;
                const int k = e / (kSN / 2), nn = 2 * (e % (kSN / 2));
                const int gn = c0 + nn;
                const int bytes = (kc + k < K) ? (gn + 1 < col_end ? 16 : (gn < col_end ? 8 : 0)) : 0;
                cpa16(Bs + k * kSLdB + nn, bytes ? X + (long long)(kc + k) * lda + gn : X, bytes);
            }
        } else {
#pragma unroll
            for (int t = 0; t < kSK * kSM / kSThreads; ++t) {
                const int e = tid + kSThreads * t;
                const int k = e / kSM, m = e % kSM;
                const bool ok = (kc + k < K) && (r0 + m < M);
                cpa8(As + k * kSLdA + m, ok ? X + (long long)(kc + k) * lda + r0 + m : X, ok);
            }
#pragma unroll
            for (int t = 0; t < kSK * kSN / kSThreads; ++t) {
                const int e = tid + kSThreads * t;
                const int k = e / kSN, nn = e % kSN;
                const bool ok = (kc + k < K) && (c0 + nn < col_end);
                cpa8(Bs + k * kSLdB + nn, ok ? X + (long long)(kc + k) * lda + c0 + nn : X, ok);
            }
        }
    };

    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int r = lane >> 2, q = lane & 3;
    stage_load(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    // interior tiles (inside the matrix, entirely on or below the diagonal) need no predicates
    const bool interior = (r0 + kSM <= M) && (c0 + kSN <= col_end) && (r0 >= c0 + kSN - 1);
    double* const cbase = A + (long long)(c0 + wn + 2 * q) * lda + r0 + wm + r;     // element (i, j, h): + (8 j + h) lda + 8 i
    double acc[4][4][2];
    if constexpr (kRed) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    } else if (interior) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][j][0] = -cbase[(long long)(8 * j) * lda + 8 * i];
                acc[i][j][1] = -cbase[(long long)(8 * j + 1) * lda + 8 * i];
            }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int gm = r0 + wm + 8 * i + r;
                const int gn = c0 + wn + 8 * j + 2 * q;
                acc[i][j][0] = (gm < M && gn < col_end && gm >= gn) ? -cbase[(long long)(8 * j) * lda + 8 * i] : 0.0;
                acc[i][j][1] = (gm < M && gn + 1 < col_end && gm >= gn + 1) ? -cbase[(long long)(8 * j + 1) * lda + 8 * i] : 0.0;
            }
    }

    int slot = 0;
    for (int g = 0; g < nchunks; ++g) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (g + 1 < nchunks) stage_load(slot ^ 1, g + 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* As = smem_syrk + (size_t)slot * kSStageDoubles;
        const double* Bs = As + kSK * kSLdA;
#pragma unroll
        for (int ks = 0; ks < kSK; ks += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(ks + q) * kSLdA + wm + 8 * i + r];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(ks + q) * kSLdB + wn + 8 * j + r];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        slot ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (interior) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double* p0 = cbase + (long long)(8 * j) * lda + 8 * i;
                if constexpr (kRed) { atomicAdd(p0, -acc[i][j][0]); atomicAdd(p0 + lda, -acc[i][j][1]); }
                else { p0[0] = -acc[i][j][0]; p0[lda] = -acc[i][j][1]; }
            }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int gm = r0 + wm + 8 * i + r;
                const int gn = c0 + wn + 8 * j + 2 * q;
                double* p0 = cbase + (long long)(8 * j) * lda + 8 * i;
                if (gm < M && gn < col_end && gm >= gn) {
                    if constexpr (kRed) atomicAdd(p0, -acc[i][j][0]);
                    else p0[0] = -acc[i][j][0];
                }
                if (gm < M && gn + 1 < col_end && gm >= gn + 1) {
                    if constexpr (kRed) atomicAdd(p0 + lda, -acc[i][j][1]);
                    else p0[lda] = -acc[i][j][1];
                }
            }
    }
}

// ---- Schur complement + z = y - Y_u mu --------------------------------------------------------------------------------------
// After nk columns of the factorisation: rows nk..n-1 hold Y_u^T = (L^-1 U)^T, rows n..n+nrhs-1 hold y^T = (L^-1 b_K)^T,
// the corner block (rows/cols nk..n-1, lower) holds -S = -(Y_u^T Y_u) and A[n + r, nk + u] holds b_u - y^T Y_u.
// Every CTA factors S redundantly (nu <= 64), solves S mu = Y_u^T y - b_u and writes z for its share of the rows.
__global__ void __launch_bounds__(256) schur_kernel(int n, int nk, const double* __restrict__ A, int lda, double* __restrict__ b,
                                                    int nrhs, int ldb, int* __restrict__ info) {
    __shared__ double S[kMaxNu * (kMaxNu + 1)];
    __shared__ double mu[kMaxNu];
    const int nu = n - nk;
    const int tid = threadIdx.x;
    const int ld = kMaxNu + 1;
    if (nu > 0) {
        for (int e = tid; e < nu * nu; e += 256) {
            const int c = e / nu, i = e - c * nu;
            if (i >= c) S[c * ld + i] = -A[(long long)(nk + c) * lda + nk + i];
        }
        __syncthreads();
        chol_smem(S, ld, nu, blockIdx.x == 0 ? info : nullptr, nk, nullptr);
    }
    for (int rr = 0; rr < nrhs; ++rr) {
        if (nu > 0) {
            __syncthreads();
            if (tid == 0) {
                // S mu = g, g_u = -(A[n + rr, nk + u]);  S = Ls Ls^T
                for (int u = 0; u < nu; ++u) {
                    double v = -A[(long long)(nk + u) * lda + n + rr];
                    for (int p = 0; p < u; ++p) v = fma(-S[p * ld + u], mu[p], v);
                    mu[u] = v / S[u * ld + u];
                }
                for (int u = nu - 1; u >= 0; --u) {
                    double v = mu[u];
                    for (int p = u + 1; p < nu; ++p) v = fma(-S[u * ld + p], mu[p], v);
                    mu[u] = v / S[u * ld + u];
                }
            }
            __syncthreads();
            if (blockIdx.x == 0 && tid < nu) b[(long long)rr * ldb + nk + tid] = mu[tid];
        }
        for (long long i = (long long)blockIdx.x * 256 + tid; i < nk; i += (long long)gridDim.x * 256) {
            const double* col = A + i * lda;
            double z = col[n + rr];
            for (int u = 0; u < nu; ++u) z = fma(-col[nk + u], mu[u], z);
            b[(long long)rr * ldb + i] = z;
        }
    }
}

// ---- backward substitution L^T x = z ---------------------------------------------------------------------------------------
// CTA `blockIdx.x` owns block row I = nblk - 1 - blockIdx.x (64 unknowns), so CTAs are dispatched in dependency order.
// x_I = L_II^-T (z_I - sum_{J > I} L[J, I]^T x_J): the columns of L[J, I] are contiguous (column-major), every warp owns 16
// columns and keeps per-lane partial sums; solution blocks are consumed as their flags appear; the diagonal solve is a
// 64 x 64 matrix-vector product with the block's L_II^-T the factorisation left in the workspace (Wt_all).
__global__ void __launch_bounds__(128) trsv_lt_kernel(int nk, const double* __restrict__ L, int lda, double* __restrict__ x,
                                                      const double* __restrict__ Wt_all, int* __restrict__ flags) {
    __shared__ double Ws[kCB * (kCB + 1)];       // Ws[k * 65 + n] = (L_II^-T)[k][n]
    __shared__ double xs[kCB];
    __shared__ double part[kCB];
    const int nblk = (nk + kCB - 1) / kCB;
    const int I = nblk - 1 - (int)blockIdx.x;
    const int c0 = I * kCB;
    const int jb = min(kCB, nk - c0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const double* Wt = Wt_all + (long long)I * kCB * kCB;
        double tmp[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) tmp[t] = Wt[tid + 128 * t];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const int e = tid + 128 * t;
            Ws[(e >> 6) * (kCB + 1) + (e & 63)] = tmp[t];
        }
    }
    const double z_own = (tid < jb) ? x[c0 + tid] : 0.0;
    double acc[16];
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) acc[cc] = 0.0;
    for (int J = nblk - 1; J > I; --J) {
        const int rb0 = J * kCB;
        const int rb = min(kCB, nk - rb0);
        // the factor is final: fetch this warp's 16 column segments before waiting for x_J
        double v0[16], v1[16];
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
            const int c = warp * 16 + cc;
            const double* col = L + (long long)(c0 + c) * lda + rb0;
            v0[cc] = (c < jb && lane < rb) ? col[lane] : 0.0;
            v1[cc] = (c < jb && lane + 32 < rb) ? col[lane + 32] : 0.0;
        }
        if (tid == 0) {
            while (atomicAdd(&flags[J], 0) == 0) { }
            __threadfence();
        }
        __syncthreads();
        if (tid < kCB) xs[tid] = (tid < rb) ? __ldcg(x + rb0 + tid) : 0.0;
        __syncthreads();
        const double xlo = xs[lane], xhi = xs[lane + 32];
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) acc[cc] = fma(v1[cc], xhi, fma(v0[cc], xlo, acc[cc]));
        __syncthreads();                         // xs is rewritten in the next iteration
    }
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) {
        double s = acc[cc];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) part[warp * 16 + cc] = s;
    }
    __syncthreads();
    if (tid < kCB) xs[tid] = (tid < jb) ? z_own - part[tid] : 0.0;       // v = z_I - sum
    __syncthreads();
    {   // x_I[k] = sum_{n >= k} Wt[k][n] v[n]: two threads per k, 32 columns each
        const int k = tid >> 1, half = tid & 1;
        double s = 0.0;
#pragma unroll 8
        for (int n = half * 32; n < half * 32 + 32; ++n) s = fma(Ws[k * (kCB + 1) + n], xs[n], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (half == 0 && k < jb) x[c0 + k] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(&flags[I], 1);
}

// ---- host side ------------------------------------------------------------------------------------------------------------
int syrk_mode() {       // 1 (default): L2 reductions, 0: load / subtract / store epilogue (GPB_CHOL_RED=0)
    static const int v = [] { const char* e = getenv("GPB_CHOL_RED"); return e ? atoi(e) : 1; }();
    return v;
}

int launch_syrk(int M, int K, int src, int col_begin, int col_end, double* A, int lda, cudaStream_t s) {
    static bool attr_sets[GPB_MAX_DEVICES] = {false};
    bool& attr_set = attr_sets[gpb_current_device()];
    if (!attr_set) {
        GPB_CHECK_CUDA(cudaFuncSetAttribute(syrk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(syrk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(syrk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(syrk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(syrk_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(syrk_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(syrk_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(syrk_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    if (col_end <= col_begin || M <= col_begin || K <= 0) return GPB_OK;
    const dim3 grid((M - col_begin + kSM - 1) / kSM, (col_end - col_begin + kSN - 1) / kSN);
    const double* X = A + (long long)src * lda;
    const bool aligned = (lda % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (col_begin % 2 == 0);
    const bool red = syrk_mode() == 1;
    if (aligned && red) syrk_kernel<true, true><<<grid, kSThreads, kSyrkSmem, s>>>(M, K, X, A, lda, col_begin, col_end);
    else if (aligned) syrk_kernel<true, false><<<grid, kSThreads, kSyrkSmem, s>>>(M, K, X, A, lda, col_begin, col_end);
    else if (red) syrk_kernel<false, true><<<grid, kSThreads, kSyrkSmem, s>>>(M, K, X, A, lda, col_begin, col_end);
    else syrk_kernel<false, false><<<grid, kSThreads, kSyrkSmem, s>>>(M, K, X, A, lda, col_begin, col_end);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

// columns [K0, K0 + W) of the factor, 64 at a time, updates confined to the block's own columns
int chol_panel_attrs() {
    static bool attr_sets[GPB_MAX_DEVICES] = {false};
    bool& attr_set = attr_sets[gpb_current_device()];
    if (!attr_set) {
        GPB_CHECK_CUDA(cudaFuncSetAttribute(trsm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsmSmem));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(trsm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsmSmem));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(potrf64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotrfSmem));
        attr_set = true;
    }
    return GPB_OK;
}

int launch_trsm(int M, int k0, int jb, double* A, int lda, const double* Wt, cudaStream_t q) {
    const int rows = M - (k0 + jb);
    if (rows <= 0) return GPB_OK;
    const bool aligned = (lda % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0) && ((k0 + jb) % 2 == 0);
    const int grid = (rows + kCB - 1) / kCB;
    if (aligned) trsm_dmma_kernel<true><<<grid, 128, kTrsmSmem, q>>>(M, k0, jb, A, lda, Wt);
    else trsm_dmma_kernel<false><<<grid, 128, kTrsmSmem, q>>>(M, k0, jb, A, lda, Wt);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

// wt_all: nblk blocks of 64 x 64 doubles (L_II^-T of every diagonal block; block index k0 / 64)
int factor_block(int M, int K0, int W, double* A, int lda, int* info, const double* diag0, double* wt_all, cudaStream_t q) {
    for (int k0 = K0; k0 < K0 + W; k0 += kCB) {
        const int jb = min(kCB, K0 + W - k0);
        double* Wt = wt_all + (long long)(k0 / kCB) * kCB * kCB;
        int rc0;
        static const double pivot_tol = [] { const char* e = getenv("GPB_CHOL_PIVOT_TOL"); return e ? atof(e) : kPivotTol; }();
        potrf64_kernel<<<1, 256, kPotrfSmem, q>>>(k0, jb, A, lda, info, diag0, Wt, pivot_tol);
        GPB_LAUNCH_CHECK();
        if ((rc0 = launch_trsm(M, k0, jb, A, lda, Wt, q))) return rc0;
        if (k0 + jb < K0 + W) {
            int rc = launch_syrk(M, jb, k0, k0 + jb, K0 + W, A, lda, q);
            if (rc) return rc;
        }
    }
    return GPB_OK;
}

// outer block width: 128 below n = 12 000, 256 above (measured: n = 6 999 8.4 vs 8.7 ms, n = 17 499 70.7 vs 66.5 ms);
// GPB_CHOL_OUTER = 64 / 128 / 192 / 256 forces one
int outer_width_chol(int nk) {
    static const int forced = [] { const char* e = getenv("GPB_CHOL_OUTER"); const int w = e ? atoi(e) : 0; return (w == 64 || w == 128 || w == 192 || w == 256) ? w : 0; }();
    if (forced) return forced;
    return nk >= 12000 ? 256 : kCOuter;
}

int chol_factor(int M, int n, int nk, double* A, int lda, int* info, const double* diag0, double* wt_all, cudaStream_t s) {
    GpbSideStream* side = gpb_side_stream();
    const bool ahead = side != nullptr && getenv("GPB_LU_NO_LOOKAHEAD") == nullptr;
    cudaStream_t ps = ahead ? side->stream : s;
    const int ow = outer_width_chol(nk);
    if (ahead) {
        GPB_CHECK_CUDA(cudaEventRecord(side->ready, s));
        GPB_CHECK_CUDA(cudaStreamWaitEvent(ps, side->ready, 0));
    }
    int rc = factor_block(M, 0, min(ow, nk), A, lda, info, diag0, wt_all, ps);
    if (rc) return rc;
    for (int K0 = 0; K0 < nk;) {
        const int W = min(ow, nk - K0);
        const int K1 = K0 + W;
        const int W1 = (K1 < nk) ? min(ow, nk - K1) : 0;
        if (ahead) {
            GPB_CHECK_CUDA(cudaEventRecord(side->done, ps));
            GPB_CHECK_CUDA(cudaStreamWaitEvent(s, side->done, 0));
        }
        if (W1 > 0) {
            if ((rc = launch_syrk(M, W, K0, K1, K1 + W1, A, lda, s))) return rc;
            if (ahead) {
                GPB_CHECK_CUDA(cudaEventRecord(side->ready, s));
                GPB_CHECK_CUDA(cudaStreamWaitEvent(ps, side->ready, 0));
            }
            if ((rc = factor_block(M, K1, W1, A, lda, info, diag0, wt_all, ps))) return rc;
            if ((rc = launch_syrk(M, W, K0, K1 + W1, n, A, lda, s))) return rc;
        } else {
            if ((rc = launch_syrk(M, W, K0, K1, n, A, lda, s))) return rc;      // the Schur corner (columns nk .. n-1)
        }
        K0 = K1;
    }
    if (ahead) {
        GPB_CHECK_CUDA(cudaEventRecord(side->done, ps));
        GPB_CHECK_CUDA(cudaStreamWaitEvent(s, side->done, 0));
    }
    return GPB_OK;
}

}  // namespace

extern "C" int gpb_sym_solve(int n, int nk, double* A, int lda, double* b, int nrhs, int ldb, int* info, void* stream) {
    GPB_REQUIRE(n > 0 && nk >= 1 && nk <= n && n - nk <= kMaxNu, "bad sizes (1 <= nk <= n, n - nk <= 64)");
    GPB_REQUIRE(A && b && nrhs >= 1 && nrhs <= 16 && ldb >= n && lda >= n + nrhs, "bad arguments (lda >= n + nrhs: the right-hand sides ride along as extra rows)");
    cudaStream_t s = (cudaStream_t)stream;
    const int M = n + nrhs;
    GpbDeviceLock lock;                                  // one enqueue at a time per device (shared side stream / events)
    zero_int_kernel<<<1, 1, 0, s>>>(info);
    GPB_LAUNCH_CHECK();
    {
        const long long total = (long long)n * nrhs;
        set_rhs_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(n, A, lda, b, nrhs, ldb);
        GPB_LAUNCH_CHECK();
    }
    // workspace: the original diagonal of K (pivot threshold) + the flags of the backward substitution
    const int nblk = (nk + kCB - 1) / kCB;
    double* ws = nullptr;
    const long long wt_doubles = (long long)nblk * kCB * kCB;
    const long long nk_pad = (nk + 1) & ~1LL;            // keeps wt_all 16-byte aligned (cudaMallocAsync aligns to >= 256 B)
    {
        int rca = chol_panel_attrs();
        if (rca) return rca;
    }
    GPB_CHECK_CUDA(gpb_malloc_async((void**)&ws, sizeof(double) * (nk_pad + wt_doubles) + sizeof(int) * nblk, s));
    double* wt_all = ws + nk_pad;
    int* flags = reinterpret_cast<int*>(ws + nk_pad + wt_doubles);
    save_diag_kernel<<<(nk + 255) / 256, 256, 0, s>>>(nk, A, lda, ws);
    GPB_LAUNCH_CHECK();
    int rc = chol_factor(M, n, nk, A, lda, info, ws, wt_all, s);
    if (rc) { cudaFreeAsync(ws, s); return rc; }
    int grid = (nk + 255) / 256;
    if (grid > 64) grid = 64;
    schur_kernel<<<grid, 256, 0, s>>>(n, nk, A, lda, b, nrhs, ldb, info);
    GPB_LAUNCH_CHECK();
    for (int r = 0; r < nrhs; ++r) {
        GPB_CHECK_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * nblk, s));
        trsv_lt_kernel<<<nblk, 128, 0, s>>>(nk, A, lda, b + (long long)r * ldb, wt_all, flags);
        GPB_LAUNCH_CHECK();
    }
    GPB_CHECK_CUDA(cudaFreeAsync(ws, s));
    return GPB_OK;
}
