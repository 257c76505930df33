// (2) Dense FP64 solve of the saddle-point system: blocked right-looking LU with partial pivoting.
//
// Engine stage replaced: "solver" with kernel_solver = 1 (direct dense solve; the reference calls LAPACK gesv
// through numpy.linalg.solve) -- SURVEY.md 8a2 row (2); reference call site gempy/API/compute_API.py:68-73.
//
// Structure (column-major, in place):
//   n <= kSmallN : one CTA, whole matrix in shared memory (the reference's example models: n = 8 ... 104).
//   otherwise, for each panel of kNB columns:
//     panel_kernel   unblocked partial-pivot LU of the tall panel (pivot search = block reduction)
//     swap_trsm      row interchanges of the panel applied to the columns right of it + U12 = L11^-1 A12
//     gemm_kernel    A22 -= L21 * U12 on the FP64 tensor cores (mma.sync m8n8k4 DMMA) -- the one dense contraction
//   Interchanges are applied LAPACK-style inside a panel and LINPACK-style across panels (columns left of a
//   panel are never permuted); gpb_lu_apply replays them panel by panel, so factor + apply are self-consistent.
//   n >= 10240: outer blocks of 256 columns (factor_outer): the block is factored panel by panel on the side stream,
//   the rest of the matrix gets one K = 256 DMMA update per block (gemm_big_kernel, C updated by L2 reductions).
#include "gpb_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int kNB = 32;          // (maximum) panel width = K of the trailing update
constexpr int kSmallN = 160;     // whole-matrix-in-smem path

// Panel width at column k0 of an n x n factorisation: a pure function of (n, k0) and the device's cluster
// configuration, evaluated identically by gpb_lu_factor (host) and by the forward substitution of gpb_lu_apply
// (device), so the interchanges are replayed over exactly the panels that produced them.
__host__ __device__ inline int panel_width_hd(int n, int k0, int cluster, unsigned long long smem_cap) {
    (void)cluster; (void)smem_cap;               // the panel width no longer depends on the device configuration
    const int jb = kNB;
    return (n - k0 < jb) ? n - k0 : jb;
}

// =====================================================================================================
// small systems: one CTA, matrix in shared memory
// =====================================================================================================
__global__ void __launch_bounds__(256) lu_small_kernel(int n, double* __restrict__ A, int lda, double* __restrict__ b,
                                                        int nrhs, int ldb, int* __restrict__ ipiv, int* __restrict__ info,
                                                        int do_solve) {
    extern __shared__ double sm[];
    const int ld = n + 1;                       // padded leading dimension
    double* S = sm;                             // n x n, column-major, ld
    __shared__ double red_v[8];
    __shared__ int red_i[8];
    __shared__ int piv_s;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < n * n; e += nt) {
        const int j = e / n, i = e - j * n;
        S[j * ld + i] = A[(long long)j * lda + i];
    }
    if (tid == 0 && info) *info = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        // pivot search in column k, rows k..n-1 (first maximum, like idamax)
        double best = -1.0;
        int bi = n;
        for (int i = k + tid; i < n; i += nt) {
            const double v = fabs(S[k * ld + i]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { red_v[tid >> 5] = best; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (nt >> 5); ++w)
                if (red_v[w] > best || (red_v[w] == best && red_i[w] < bi)) { best = red_v[w]; bi = red_i[w]; }
            piv_s = bi;
            ipiv[k] = bi;
            if (best == 0.0 && info && *info == 0) *info = k + 1;
        }
        __syncthreads();
        const int p = piv_s;
        if (p != k) {
            for (int j = (k / kNB) * kNB + tid; j < n; j += nt) {     // LINPACK-style across panels
                const double t = S[j * ld + k];
                S[j * ld + k] = S[j * ld + p];
                S[j * ld + p] = t;
            }
        }
        __syncthreads();
        const double pv = S[k * ld + k];
        const double inv = pv != 0.0 ? 1.0 / pv : 0.0;
        for (int i = k + 1 + tid; i < n; i += nt) S[k * ld + i] *= inv;
        __syncthreads();
        const int rem = n - k - 1;
        for (int e = tid; e < rem * rem; e += nt) {
            const int jj = e / rem, ii = e - jj * rem;
            const int i = k + 1 + ii, j = k + 1 + jj;
            S[j * ld + i] = fma(-S[k * ld + i], S[j * ld + k], S[j * ld + i]);
        }
        __syncthreads();
    }
    for (int e = tid; e < n * n; e += nt) {
        const int j = e / n, i = e - j * n;
        A[(long long)j * lda + i] = S[j * ld + i];
    }
    if (!do_solve) return;
    // solve for each right-hand side (one warp per rhs would be enough; systems are tiny)
    for (int r = 0; r < nrhs; ++r) {
        double* x = b + (long long)r * ldb;
        __syncthreads();
        __shared__ double xs[kSmallN];
        for (int i = tid; i < n; i += nt) xs[i] = x[i];
        __syncthreads();
        for (int k = 0; k < n; ++k) {               // forward, unit lower; swaps replayed panel by panel
            if (k % kNB == 0) {
                if (tid == 0) {
                    const int ke = min(k + kNB, n);
                    for (int kk = k; kk < ke; ++kk) {
                        const int p = ipiv[kk];
                        if (p != kk) { const double t = xs[kk]; xs[kk] = xs[p]; xs[p] = t; }
                    }
                }
                __syncthreads();
            }
            const double xk = xs[k];
            for (int i = k + 1 + tid; i < n; i += nt) xs[i] = fma(-S[k * ld + i], xk, xs[i]);
            __syncthreads();
        }
        for (int k = n - 1; k >= 0; --k) {          // backward
            if (tid == 0) xs[k] /= S[k * ld + k];
            __syncthreads();
            const double xk = xs[k];
            for (int i = tid; i < k; i += nt) xs[i] = fma(-S[k * ld + i], xk, xs[i]);
            __syncthreads();
        }
        for (int i = tid; i < n; i += nt) x[i] = xs[i];
    }
}

// =====================================================================================================
// blocked path
// =====================================================================================================
// Unblocked partial-pivot LU of the panel A[k0:n, k0:k0+jb]; one CTA of 1024 threads, panel in global/L2.
__global__ void __launch_bounds__(1024) panel_kernel(int n, int k0, int jb, double* __restrict__ A, int lda,
                                                      int* __restrict__ ipiv, int* __restrict__ info) {
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    __shared__ int piv_s;
    __shared__ double prow[kNB];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = 0; j < jb; ++j) {
        const int kj = k0 + j;
        double* col = A + (long long)kj * lda;
        double best = -1.0;
        int bi = n;
        for (int i = kj + tid; i < n; i += nt) {
            const double v = fabs(col[i]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { red_v[tid >> 5] = best; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (nt >> 5); ++w)
                if (red_v[w] > best || (red_v[w] == best && red_i[w] < bi)) { best = red_v[w]; bi = red_i[w]; }
            piv_s = bi;
            ipiv[kj] = bi;
            if (best == 0.0 && info && *info == 0) *info = kj + 1;
        }
        __syncthreads();
        const int p = piv_s;
        // swap rows kj <-> p inside the panel and stage the pivot row
        if (tid < jb) {
            double* c = A + (long long)(k0 + tid) * lda;
            const double vp = c[p];
            if (p != kj) { c[p] = c[kj]; c[kj] = vp; }
            prow[tid] = vp;
        }
        __syncthreads();
        const double pv = prow[j];
        const double inv = pv != 0.0 ? 1.0 / pv : 0.0;
        const int nc = jb - j - 1;
        for (int i = kj + 1 + tid; i < n; i += nt) {
            const double l = col[i] * inv;
            col[i] = l;
            for (int c = 0; c < nc; ++c) {
                double* a = A + (long long)(kj + 1 + c) * lda + i;
                *a = fma(-l, prow[j + 1 + c], *a);
            }
        }
        __syncthreads();
    }
}


// -----------------------------------------------------------------------------------------------------
// Cluster panel: the rows of the tall panel are distributed over the shared memory of the CTAs of one
// thread-block cluster (up to 16 SMs).  Per column: local arg-max -> candidates exchanged through distributed
// shared memory -> cluster barrier -> the pivot row is broadcast (and the displaced top row sent to the pivot's
// owner) through DSMEM -> cluster barrier -> CTA-local rank-1 update out of shared memory.  Two cluster
// barriers per column replace the grid-wide synchronisation a multi-CTA panel would otherwise need.
// -----------------------------------------------------------------------------------------------------
constexpr int kPanelThreads = 512;
constexpr int kMaxCluster = 16;

// Per-column exchange area (lives in every CTA's shared memory, written remotely through DSMEM).
struct PanelExchange {
    double cand_v[kMaxCluster];            // |pivot candidate| of each CTA (-1: no active row)
    int cand_i[kMaxCluster];               // its global row
    double cand_row[kMaxCluster][kNB];     // the candidate's row of the panel
    double top_row[kNB];                   // row k0 + j (the row the pivot is exchanged with)
};

__device__ __forceinline__ void better(double& bv, int& bi, double ov, int oi) {
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
}

__global__ void __launch_bounds__(kPanelThreads, 1)
panel_cluster_kernel(int n, int k0, int jb, double* __restrict__ A, int lda, int* __restrict__ ipiv, int* __restrict__ info,
                     int R, int ldp, int in_smem) {
    extern __shared__ double P_smem[];             // [jb][ldp] : this CTA's rows of the panel, column-major
    __shared__ PanelExchange ex[2];                // double buffered by column parity
    __shared__ double red_v[kPanelThreads / 32];
    __shared__ int red_i[kPanelThreads / 32];
    __shared__ double loc_v, inv_s;
    __shared__ int loc_i, win_s, piv_s;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = k0 + rank * R;
    const int nrows = max(0, min(n, r0 + R) - r0);
    // panels too tall for shared memory (n > ~12 800 at 16 CTAs) are factored in place: same indexing with ldp = lda;
    // the rows of a CTA are private to it, so only the CTA's own L1/L2 path is involved
    double* const P = in_smem ? P_smem : A + (long long)k0 * lda + r0;

    if (in_smem) {
        for (int e = tid; e < nrows * jb; e += kPanelThreads) {
            const int c = e / nrows, i = e - c * nrows;
            P[c * ldp + i] = A[(long long)(k0 + c) * lda + r0 + i];
        }
    }
    __syncthreads();
    // candidate for column 0
    double best = -1.0;
    int bi = n;
    for (int i = tid; i < nrows; i += kPanelThreads) {
        const double v = fabs(P[i]);
        if (v > best) { best = v; bi = r0 + i; }
    }
    cluster.sync();                                // every CTA of the cluster is resident before any DSMEM access

    for (int j = 0; j < jb; ++j) {
        const int kj = k0 + j;
        const int buf = j & 1;
        // ---- (1) CTA-wide candidate
        for (int o = 16; o > 0; o >>= 1)
            better(best, bi, __shfl_down_sync(0xffffffffu, best, o), __shfl_down_sync(0xffffffffu, bi, o));
        if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            double v = lane < kPanelThreads / 32 ? red_v[lane] : -1.0;
            int ii = lane < kPanelThreads / 32 ? red_i[lane] : n;
            for (int o = 8; o > 0; o >>= 1)
                better(v, ii, __shfl_down_sync(0xffffffffu, v, o), __shfl_down_sync(0xffffffffu, ii, o));
            if (lane == 0) { loc_v = v; loc_i = ii; }
        }
        __syncthreads();
        const double my_v = loc_v;
        const int my_i = loc_i;
        // ---- (2) publish candidate (+ its row) and, from its owner, the top row, to every CTA
        const int owner_t = (kj - k0) / R;
        for (int e = tid; e < C * jb; e += kPanelThreads) {
            const int dst = e / jb, c = e - dst * jb;
            PanelExchange* rx = cluster.map_shared_rank(&ex[buf], dst);
            rx->cand_row[rank][c] = (my_i < n) ? P[c * ldp + (my_i - r0)] : 0.0;
            if (rank == owner_t) rx->top_row[c] = P[c * ldp + (kj - r0)];
        }
        if (tid < C) {
            PanelExchange* rx = cluster.map_shared_rank(&ex[buf], tid);
            rx->cand_v[rank] = my_v;
            rx->cand_i[rank] = my_i;
        }
        cluster.sync();                            // the only cluster barrier of this column
        // ---- (3) one warp elects the pivot (every CTA arrives at the same answer) and inverts it
        if (warp == 0) {
            double gv = lane < C ? ex[buf].cand_v[lane] : -1.0;
            int gp = lane < C ? ex[buf].cand_i[lane] : n;
            int gw = lane;
            for (int o = 8; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, gv, o);
                const int oi = __shfl_down_sync(0xffffffffu, gp, o);
                const int ow = __shfl_down_sync(0xffffffffu, gw, o);
                if (ov > gv || (ov == gv && oi < gp)) { gv = ov; gp = oi; gw = ow; }
            }
            if (lane == 0) {
                win_s = gw;
                piv_s = gp;
                const double pvv = ex[buf].cand_row[gw][j];
                inv_s = pvv != 0.0 ? 1.0 / pvv : 0.0;
                if (rank == 0) {
                    ipiv[kj] = gp;
                    if (gv == 0.0 && info && *info == 0) *info = kj + 1;
                }
            }
        }
        __syncthreads();
        const int p = piv_s;
        const double* prow = ex[buf].cand_row[win_s];
        const int owner_p = (p - k0) / R;
        if (p != kj && tid < jb) {
            if (rank == owner_t) P[tid * ldp + (kj - r0)] = prow[tid];
            if (rank == owner_p) P[tid * ldp + (p - r0)] = ex[buf].top_row[tid];
        }
        __syncthreads();
        // ---- (4) rank-1 update of this CTA's rows, fused with the candidate search of column j + 1
        const double inv = inv_s;
        best = -1.0;
        bi = n;
        for (int i = tid; i < nrows; i += kPanelThreads) {
            const int gi = r0 + i;
            if (gi > kj) {
                const double l = P[j * ldp + i] * inv;
                P[j * ldp + i] = l;
                // chunks of 8 columns: all loads first, so the shared-memory latency is paid once per chunk
                for (int c0 = j + 1; c0 < jb; c0 += 8) {
                    double v[8], pr[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int c = c0 + u;
                        v[u] = (c < jb) ? P[c * ldp + i] : 0.0;
                        pr[u] = (c < jb) ? prow[c] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = fma(-l, pr[u], v[u]);
                    if (c0 == j + 1) {
                        const double a1 = fabs(v[0]);
                        if (a1 > best) { best = a1; bi = gi; }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (c0 + u < jb) P[(c0 + u) * ldp + i] = v[u];
                }
            }
        }
        __syncthreads();
    }
    if (in_smem) {
        for (int e = tid; e < nrows * jb; e += kPanelThreads) {
            const int c = e / nrows, i = e - c * nrows;
            A[(long long)(k0 + c) * lda + r0 + i] = P[c * ldp + i];
        }
    }
    cluster.sync();                                // no CTA exits while a neighbour may still address its smem
}

// Apply the panel's interchanges to the columns right of the panel and solve U12 = L11^-1 A12.
// One CTA handles 64 columns.
// Blocks beyond the matrix' own column blocks work on the right-hand sides B (n x nrhs, ldb): the forward
// substitution of gpb_lu_solve is carried along with the factorisation.
// col_begin/col_end bound the matrix columns handled (the outer-blocked path restricts them); do_swap = 0 skips the
// interchanges (already applied by laswp_kernel).
__global__ void __launch_bounds__(256) swap_trsm_kernel(int n, int k0, int jb, double* __restrict__ A, int lda,
                                                        const int* __restrict__ ipiv, int n_mat_blocks,
                                                        double* __restrict__ B, int ldb, int nrhs,
                                                        int col_begin, int col_end, int do_swap) {
    __shared__ double L[kNB][kNB + 1];
    __shared__ double T[64][kNB + 1];
    __shared__ int piv[kNB];
    const int tid = threadIdx.x;
    const bool on_rhs = (int)blockIdx.x >= n_mat_blocks;
    const int c0 = on_rhs ? ((int)blockIdx.x - n_mat_blocks) * 64 : col_begin + blockIdx.x * 64;
    const int ncol = on_rhs ? min(64, nrhs - c0) : min(64, col_end - c0);
    double* const M = on_rhs ? B : A;              // the columns this block transforms
    const int ldm = on_rhs ? ldb : lda;
    for (int e = tid; e < jb * jb; e += 256) {
        const int j = e / jb, i = e - j * jb;
        L[i][j] = A[(long long)(k0 + j) * lda + k0 + i];
    }
    if (tid < jb) piv[tid] = ipiv[k0 + tid];
    __syncthreads();
    // interchanges: one thread per column, sequential over the panel's pivots
    if (do_swap && tid < ncol) {
        double* c = M + (long long)(c0 + tid) * ldm;
        for (int j = 0; j < jb; ++j) {
            const int p = piv[j];
            if (p != k0 + j) { const double t = c[k0 + j]; c[k0 + j] = c[p]; c[p] = t; }
        }
    }
    __syncthreads();
    for (int e = tid; e < ncol * jb; e += 256) {
        const int c = e / jb, i = e - c * jb;
        T[c][i] = M[(long long)(c0 + c) * ldm + k0 + i];
    }
    __syncthreads();
    // forward substitution with four threads per column: thread (c, part) owns rows i = part, part + 4, ... of column c;
    // after step k the value T[c][k] is final.  (One thread per column took 22 us per call -- 496 dependent
    // shared-memory FMAs -- and sat on the critical path of every panel.)
    {
        const int c = tid >> 2, part = tid & 3;
        for (int k = 0; k < jb; ++k) {
            if (c < ncol) {
                const double xk = T[c][k];
                for (int i = k + 1 + part; i < jb; i += 4) T[c][i] = fma(-L[i][k], xk, T[c][i]);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < ncol * jb; e += 256) {
        const int c = e / jb, i = e - c * jb;
        M[(long long)(c0 + c) * ldm + k0 + i] = T[c][i];
    }
}

// Right-hand sides: B[k0+jb:, :] -= L21 * B[k0:k0+jb, :]   (the gemv twin of the trailing update)
__global__ void __launch_bounds__(256) rhs_update_kernel(int n, int k0, int jb, const double* __restrict__ A, int lda,
                                                         double* __restrict__ B, int ldb, int nrhs) {
    __shared__ double xs[kNB];
    const int i = k0 + jb + blockIdx.x * 256 + threadIdx.x;
    for (int r = 0; r < nrhs; ++r) {
        double* x = B + (long long)r * ldb;
        __syncthreads();
        if (threadIdx.x < jb) xs[threadIdx.x] = x[k0 + threadIdx.x];
        __syncthreads();
        if (i < n) {
            double v = x[i];
            for (int c = 0; c < jb; ++c) v = fma(-A[(long long)(k0 + c) * lda + i], xs[c], v);
            x[i] = v;
        }
    }
}

// Backward substitution U x = y with one CTA per block of kNB rows.  CTA `blockIdx.x` owns the block row
// nblk - 1 - blockIdx.x, so CTAs are scheduled in dependency order; a CTA consumes the solution blocks below it as
// their flags appear (release/acquire through global memory) and publishes its own block when done.
__global__ void __launch_bounds__(128) trsv_upper_kernel(int n, const double* __restrict__ LU, int lda, double* __restrict__ x,
                                                         int* __restrict__ flags) {
    __shared__ double D[kNB][kNB + 1];
    __shared__ double xk[kNB];
    __shared__ double part[4][kNB];
    const int nblk = (n + kNB - 1) / kNB;
    const int j = nblk - 1 - (int)blockIdx.x;
    const int r0 = j * kNB;
    const int jb = min(kNB, n - r0);
    const int tid = threadIdx.x, row = tid & 31, q = tid >> 5;
    for (int e = tid; e < jb * jb; e += 128) {
        const int c = e / jb, i = e - c * jb;
        D[i][c] = LU[(long long)(r0 + c) * lda + r0 + i];
    }
    double acc = 0.0;
    for (int k = nblk - 1; k > j; --k) {
        const int c0 = k * kNB;
        const int kb = min(kNB, n - c0);
        if (tid == 0) {
            while (atomicAdd(&flags[k], 0) == 0) { __nanosleep(20); }
            __threadfence();
        }
        __syncthreads();
        if (tid < kb) xk[tid] = __ldcg(x + c0 + tid);
        __syncthreads();
        if (row < jb) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const int c = q * 8 + cc;
                if (c < kb) acc = fma(LU[(long long)(c0 + c) * lda + r0 + row], xk[c], acc);
            }
        }
    }
    part[q][row] = acc;
    __syncthreads();
    if (tid < 32) {
        double v = 0.0;
        if (tid < jb) v = x[r0 + tid] - (part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid]);
        for (int c = jb - 1; c >= 0; --c) {
            if (tid == c) v /= D[c][c];
            const double xc = __shfl_sync(0xffffffffu, v, c);
            if (tid < c) v = fma(-D[tid][c], xc, v);
        }
        if (tid < jb) x[r0 + tid] = v;
        __threadfence();
        __syncwarp();
        if (tid == 0) atomicExch(&flags[j], 1);
    }
}

// ---- DMMA trailing update: C[M x N] -= Ap[M x K] * Bp[K x N], K = jb <= 32 ---------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int kGM = 64, kGN = 64;       // CTA tile
constexpr int kLdA = kGM + 4;           // As[k][m]: ld = 4 mod 16 doubles -> each half-warp of a 64-bit fragment load hits 16 distinct bank pairs
constexpr int kLdB = kNB + 4;           // Bs[n][k]: ld = 4 mod 16 doubles

__global__ void __launch_bounds__(256) gemm_kernel(int M, int N, int K, const double* __restrict__ Ap, const double* __restrict__ Bp,
                                                   double* __restrict__ C, int lda) {
    __shared__ double As[kNB * kLdA];
    __shared__ double Bs[kGN * kLdB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * kGM, n0 = blockIdx.y * kGN;
    // stage A: columns k are contiguous in m
    for (int e = tid; e < K * kGM; e += 256) {
        const int k = e / kGM, m = e - k * kGM;
        As[k * kLdA + m] = (m0 + m < M) ? -Ap[(long long)k * lda + m0 + m] : 0.0;     // negated: C += (-A) B
    }
    for (int e = tid; e < kGN * K; e += 256) {
        const int nn = e / K, k = e - nn * K;
        Bs[nn * kLdB + k] = (n0 + nn < N) ? Bp[(long long)(n0 + nn) * lda + k] : 0.0;
    }
    __syncthreads();
    // 8 warps: 2 along M (32 rows each) x 4 along N (16 cols each); warp tile 32 x 16 = 4 x 2 m8n8 tiles
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
    const int r = lane >> 2, q = lane & 3;
    constexpr int NJ = 2;
    double acc[4][NJ][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int gm = m0 + wm + 8 * i + r;
            const int gn = n0 + wn + 8 * j + 2 * q;
            acc[i][j][0] = (gm < M && gn < N) ? C[(long long)gn * lda + gm] : 0.0;
            acc[i][j][1] = (gm < M && gn + 1 < N) ? C[(long long)(gn + 1) * lda + gm] : 0.0;
        }
    for (int ks = 0; ks < K; ks += 4) {
        double a[4], b[NJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[(ks + q) * kLdA + wm + 8 * i + r];
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[j] = Bs[(wn + 8 * j + r) * kLdB + ks + q];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int gm = m0 + wm + 8 * i + r;
            const int gn = n0 + wn + 8 * j + 2 * q;
            if (gm < M && gn < N) C[(long long)gn * lda + gm] = acc[i][j][0];
            if (gm < M && gn + 1 < N) C[(long long)(gn + 1) * lda + gm] = acc[i][j][1];
        }
}

// ---- helpers of the outer-blocked factorisation (large n) ----------------------------------------------
// Interchanges of columns [k0, k0+jb) applied to the columns [col_begin, col_end) of A (one thread per column) and,
// by the blocks beyond n_mat_blocks, to the right-hand sides.  npiv <= kOuter.
constexpr int kOuter = 256;      // outer block width = K of the big trailing update (512 measured slower: 1.23 s vs 1.18 s at n = 35k)
__global__ void __launch_bounds__(256) laswp_kernel(int k0, int npiv, double* __restrict__ A, int lda, const int* __restrict__ ipiv,
                                                    int col_begin, int col_end, int n_mat_blocks,
                                                    double* __restrict__ B, int ldb, int nrhs) {
    __shared__ int piv[kOuter];
    const int tid = threadIdx.x;
    for (int j = tid; j < npiv; j += 256) piv[j] = ipiv[k0 + j];
    __syncthreads();
    const bool on_rhs = (int)blockIdx.x >= n_mat_blocks;
    const int c = on_rhs ? ((int)blockIdx.x - n_mat_blocks) * 256 + tid : col_begin + blockIdx.x * 256 + tid;
    if (c >= (on_rhs ? nrhs : col_end)) return;
    double* col = on_rhs ? B + (long long)c * ldb : A + (long long)c * lda;
    for (int j = 0; j < npiv; ++j) {
        const int p = piv[j];
        if (p != k0 + j) { const double t = col[k0 + j]; col[k0 + j] = col[p]; col[p] = t; }
    }
}

// ---- big trailing update: C[M x N] -= Ap[M x K] * Bp[K x N] for K up to kOuter --------------------------
// 128 x 64 CTA tile, 8 warps of 32 x 32, two CTAs per SM, K streamed in chunks of 16 through a 3-stage cp.async ring (8-byte copies: the
// leading dimension n is odd more often than not, so columns are only 8-byte aligned); accumulators in registers.
// C is never loaded by the SM: the epilogue sends -(A B) to the L2 as fire-and-forget FP64 reductions (RED.ADD.F64).
// Every element receives exactly one reduction per launch, so the result is the correctly rounded C - A B, bit for bit
// what a load / subtract / store epilogue gives -- but no warp ever waits for C (ncu on the load-modify-store variants:
// 37-60 % of the CTA's lifetime went into waiting for the 128 KB C tile with one CTA per SM).
constexpr int kBM = 128, kBN = 64, kBK = 32, kBStages = 2;
constexpr int kBTilesPerCta = 1;
constexpr int kBThreads = 256;
constexpr int kBLdA = kBM + 4;               // As[k][m], ld = 4 mod 16 doubles (conflict-free per half-warp)
constexpr int kBLdB = kBK + 4;               // Bs[n][k]
constexpr int kBStageDoubles = kBK * kBLdA + kBN * kBLdB;
constexpr size_t kBigGemmSmem = (size_t)kBStages * kBStageDoubles * sizeof(double);

__device__ __forceinline__ void cp_async8(double* dst, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 8 : 0;              // src-size 0 -> the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void cp_async16(double* dst, const double* src, int bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);          // bytes in {0, 8, 16}: the rest is zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// kAligned16: every operand column starts on a 16-byte boundary (even leading dimension, 16-byte aligned base), so the
// chunks move as 16-byte copies that bypass L1 -- half the copy instructions of the 8-byte path.
template <bool kAligned16>
__global__ void __launch_bounds__(kBThreads, 2) gemm_big_kernel(int M, int N, int K, const double* __restrict__ Ap,
                                                          const double* __restrict__ Bp, double* __restrict__ C, int lda, int tiles_per_cta) {
    extern __shared__ __align__(16) double smem_big[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = (K + kBK - 1) / kBK;
    // this CTA owns tiles_per_cta consecutive tiles (m fastest) and streams their K chunks through ONE continuous
    // cp.async ring, so the first chunks of the next tile are already in flight while the current tile's reductions
    // leave the SM.  Not fully persistent on purpose: CTAs that retire every few hundred microseconds are what lets
    // the look-ahead panel (a 16-SM cluster on the high-priority stream) get its SMs.
    const int tiles_m = (M + kBM - 1) / kBM, tiles_n = (N + kBN - 1) / kBN;
    const int ntiles = tiles_m * tiles_n;
    const int first_tile = (int)blockIdx.x * tiles_per_cta;
    const int my_tiles = min(tiles_per_cta, ntiles - first_tile);
    const int total = my_tiles * nchunks;

    // producer cursor (tile, chunk) of the next stage_load
    int p_m0 = (first_tile % tiles_m) * kBM, p_n0 = (first_tile / tiles_m) * kBN, p_tile = first_tile, p_chunk = 0;
    auto stage_load = [&](int slot) {
        double* As = smem_big + (size_t)slot * kBStageDoubles;
        double* Bs = As + kBK * kBLdA;
        const int kc = p_chunk * kBK;
        if constexpr (kAligned16) {
            // A chunk: kBK columns (k) x kBM rows (m), two rows per copy
#pragma unroll
            for (int t = 0; t < kBK * kBM / 2 / kBThreads; ++t) {
                const int e = tid + kBThreads * t;
                const int k = e / (kBM / 2), m = 2 * (e % (kBM / 2));
                const int gm = p_m0 + m;
                const int bytes = (kc + k < K) ? (gm + 1 < M ? 16 : (gm < M ? 8 : 0)) : 0;
                cp_async16(As + k * kBLdA + m, bytes ? Ap + (long long)(kc + k) * lda + gm : Ap, bytes);
            }
            // B chunk: kBN columns (n) x kBK rows (k), two rows per copy
#pragma unroll
            for (int t = 0; t < kBN * kBK / 2 / kBThreads; ++t) {
                const int e = tid + kBThreads * t;
                const int nn = e / (kBK / 2), k = 2 * (e % (kBK / 2));
                const int bytes = (p_n0 + nn < N) ? (kc + k + 1 < K ? 16 : (kc + k < K ? 8 : 0)) : 0;
                cp_async16(Bs + nn * kBLdB + k, bytes ? Bp + (long long)(p_n0 + nn) * lda + kc + k : Bp, bytes);
            }
        } else {
            // A chunk: kBK columns (k) x kBM rows (m), m contiguous in memory
#pragma unroll
            for (int t = 0; t < kBK * kBM / kBThreads; ++t) {
                const int e = tid + kBThreads * t;
                const int k = e / kBM, m = e % kBM;
                const bool ok = (kc + k < K) && (p_m0 + m < M);
                cp_async8(As + k * kBLdA + m, ok ? Ap + (long long)(kc + k) * lda + p_m0 + m : Ap, ok);
            }
            // B chunk: kBN columns (n) x kBK rows (k), k contiguous in memory
#pragma unroll
            for (int t = 0; t < kBN * kBK / kBThreads; ++t) {
                const int e = tid + kBThreads * t;
                const int nn = e / kBK, k = e % kBK;
                const bool ok = (kc + k < K) && (p_n0 + nn < N);
                cp_async8(Bs + nn * kBLdB + k, ok ? Bp + (long long)(p_n0 + nn) * lda + kc + k : Bp, ok);
            }
        }
        if (++p_chunk == nchunks) {
            p_chunk = 0;
            p_tile += 1;
            p_m0 = (p_tile % tiles_m) * kBM;
            p_n0 = (p_tile / tiles_m) * kBN;
        }
    };

    // 8 warps: 4 along M x 2 along N, warp tile 32 x 32 = 4 x 4 m8n8 tiles; two CTAs per SM so that one CTA's
    // reductions / first operand loads overlap the other's DMMA loop
    constexpr int TI = 4;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int r = lane >> 2, q = lane & 3;
    double acc[TI][4][2];
#pragma unroll
    for (int st = 0; st < kBStages - 1; ++st) {
        if (st < total) stage_load(st);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    int c_tile = first_tile, c_chunk = 0, slot = 0;
    for (int g = 0; g < total; ++g) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kBStages - 2) : "memory");
        __syncthreads();
        // prefetch chunk g + stages - 1 into the slot that was consumed at iteration g - 1
        if (g + kBStages - 1 < total) stage_load(slot == 0 ? kBStages - 1 : slot - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* As = smem_big + (size_t)slot * kBStageDoubles;
        const double* Bs = As + kBK * kBLdA;
#pragma unroll
        for (int ks = 0; ks < kBK; ks += 4) {
            double a[TI], b[4];
#pragma unroll
            for (int i = 0; i < TI; ++i) a[i] = As[(ks + q) * kBLdA + wm + 8 * i + r];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(wn + 8 * j + r) * kBLdB + ks + q];
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (++slot == kBStages) slot = 0;
        if (++c_chunk == nchunks) {                  // tile finished: C -= acc as reductions, restart the accumulators
            const int m0 = (c_tile % tiles_m) * kBM, n0 = (c_tile / tiles_m) * kBN;
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int gm = m0 + wm + 8 * i + r;
                    const int gn = n0 + wn + 8 * j + 2 * q;
                    if (gm < M && gn < N) atomicAdd(C + (long long)gn * lda + gm, -acc[i][j][0]);
                    if (gm < M && gn + 1 < N) atomicAdd(C + (long long)(gn + 1) * lda + gm, -acc[i][j][1]);
                    acc[i][j][0] = acc[i][j][1] = 0.0;
                }
            c_chunk = 0;
            c_tile += 1;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- triangular solves with the blocked factors (one CTA; panel-by-panel) --------------------------------
__global__ void __launch_bounds__(1024) apply_kernel(int n, const double* __restrict__ LU, int lda, const int* __restrict__ ipiv,
                                                      double* __restrict__ b, int nrhs, int ldb, int outer) {
    __shared__ double xs[kNB];
    __shared__ double D[kNB][kNB + 1];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int r = 0; r < nrhs; ++r) {
        double* x = b + (long long)r * ldb;
        // forward: P, L  (same schedule as the factorisation: interchanges are replayed per outer block -- which is
        // one kNB panel unless the outer-blocked path factored the matrix -- then the block's panels are solved)
        const int ob = outer > 0 ? outer : kNB;
        for (int K0 = 0; K0 < n; K0 += ob) {
            const int W = min(ob, n - K0);
            __syncthreads();
            if (tid == 0) {
                for (int j = 0; j < W; ++j) {
                    const int p = ipiv[K0 + j];
                    if (p != K0 + j) { const double t = x[K0 + j]; x[K0 + j] = x[p]; x[p] = t; }
                }
            }
            for (int k0 = K0; k0 < K0 + W; k0 += kNB) {
                const int jb = min(kNB, K0 + W - k0);
                __syncthreads();
                for (int e = tid; e < jb * jb; e += nt) {
                    const int c = e / jb, i = e - c * jb;
                    D[i][c] = LU[(long long)(k0 + c) * lda + k0 + i];
                }
                __syncthreads();
                if (tid < 32) {
                    double v = (tid < jb) ? x[k0 + tid] : 0.0;
                    for (int c = 0; c < jb; ++c) {
                        const double xc = __shfl_sync(0xffffffffu, v, c);
                        if (tid > c && tid < jb) v = fma(-D[tid][c], xc, v);
                    }
                    if (tid < jb) { xs[tid] = v; x[k0 + tid] = v; }
                }
                __syncthreads();
                for (int i = k0 + jb + tid; i < n; i += nt) {
                    double v = x[i];
                    for (int c = 0; c < jb; ++c) v = fma(-LU[(long long)(k0 + c) * lda + i], xs[c], v);
                    x[i] = v;
                }
            }
        }
        // backward: U
        const int last = ((n - 1) / kNB) * kNB;
        for (int k0 = last; k0 >= 0; k0 -= kNB) {
            const int jb = min(kNB, n - k0);
            __syncthreads();
            for (int e = tid; e < jb * jb; e += nt) {
                const int c = e / jb, i = e - c * jb;
                D[i][c] = LU[(long long)(k0 + c) * lda + k0 + i];
            }
            __syncthreads();
            if (tid < 32) {
                double v = (tid < jb) ? x[k0 + tid] : 0.0;
                for (int c = jb - 1; c >= 0; --c) {
                    if (tid == c) v /= D[c][c];
                    const double xc = __shfl_sync(0xffffffffu, v, c);
                    if (tid < c) v = fma(-D[tid][c], xc, v);
                }
                if (tid < jb) { xs[tid] = v; x[k0 + tid] = v; }
            }
            __syncthreads();
            for (int i = tid; i < k0; i += nt) {
                double v = x[i];
                for (int c = 0; c < jb; ++c) v = fma(-LU[(long long)(k0 + c) * lda + i], xs[c], v);
                x[i] = v;
            }
        }
        __syncthreads();
    }
}

__global__ void zero_info_kernel(int* info) { if (info) *info = 0; }

struct PanelConfig {
    int cluster = 0;          // CTAs per cluster (0: cluster panel unavailable -> single-CTA panel)
    size_t smem_cap = 0;      // usable dynamic shared memory per CTA
};

const PanelConfig& panel_config() {
    static PanelConfig cfgs[GPB_MAX_DEVICES];
    static bool dones[GPB_MAX_DEVICES] = {false};
    const int dev = gpb_current_device();
    PanelConfig& cfg = cfgs[dev];
    if (dones[dev]) return cfg;
    dones[dev] = true;
    const size_t want = 200 * 1024;
    if (cudaFuncSetAttribute(panel_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want) != cudaSuccess) {
        cudaGetLastError();
        return cfg;
    }
    cudaFuncSetAttribute(panel_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaGetLastError();
    for (int c : {16, 8, 4, 2}) {
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(c);
        lc.blockDim = dim3(kPanelThreads);
        lc.dynamicSmemBytes = want;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, panel_cluster_kernel, &lc) == cudaSuccess && nclusters >= 1) {
            cfg.cluster = c;
            cfg.smem_cap = want;
            break;
        }
        cudaGetLastError();
    }
    return cfg;
}

// Panel width at column k0 of an n x n factorisation.  A pure function of (n, k0) and the device's cluster
// configuration, so that gpb_lu_apply replays exactly the panels gpb_lu_factor used.
int panel_width(int n, int k0) {
    const PanelConfig& cfg = panel_config();
    return panel_width_hd(n, k0, cfg.cluster, (unsigned long long)cfg.smem_cap);
}

int panel_rows_cap() {
    static const int v = [] { const char* e = getenv("GPB_LU_PANEL_ROWS_CAP"); return e ? atoi(e) : 256; }();
    return v;
}

int launch_panel(int n, int k0, int jb, double* A, int lda, int* ipiv, int* info, cudaStream_t s) {
    const PanelConfig& cfg = panel_config();
    if (cfg.cluster == 0) {
        panel_kernel<<<1, 1024, 0, s>>>(n, k0, jb, A, lda, ipiv, info);
        GPB_LAUNCH_CHECK();
        return GPB_OK;
    }
    const int m = n - k0;
    // short panels run on a smaller cluster (cheaper barrier per column) as long as a CTA keeps at most
    // panel_rows_cap() rows
    int cl = cfg.cluster;
    while (cl > 2 && (m + cl / 2 - 1) / (cl / 2) <= panel_rows_cap()) cl /= 2;
    int R = (m + cl - 1) / cl;
    if (R < 1) R = 1;
    const bool in_smem = (size_t)(R + 2) * jb * sizeof(double) <= cfg.smem_cap;
    const int ldp = in_smem ? ((R + 1) & ~1) : lda;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(cl);
    lc.blockDim = dim3(kPanelThreads);
    lc.dynamicSmemBytes = in_smem ? (size_t)ldp * jb * sizeof(double) : 0;
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    GPB_CHECK_CUDA(cudaLaunchKernelEx(&lc, panel_cluster_kernel, n, k0, jb, A, lda, ipiv, info, R, ldp, in_smem ? 1 : 0));
    ++g_gpb_launches;
    return GPB_OK;
}

struct LookAhead {
    cudaStream_t panel_stream = nullptr;
    cudaEvent_t ready = nullptr, panel_done = nullptr;
    bool ok = false;
};

// the side stream and its events are per device (gpb_side_stream); callers hold a GpbDeviceLock while enqueuing
LookAhead look_ahead() {
    LookAhead la;
    static const bool disabled = getenv("GPB_LU_NO_LOOKAHEAD") != nullptr;      // diagnostic: everything on the caller's stream
    if (disabled) return la;
    if (GpbSideStream* sd = gpb_side_stream()) {
        la.panel_stream = sd->stream;
        la.ready = sd->ready;
        la.panel_done = sd->done;
        la.ok = true;
    }
    return la;
}

void launch_gemm(int n, int k0, int jb, int col_begin, int col_end, double* A, int lda, cudaStream_t s, int row_end = -1) {
    const int M = (row_end < 0 ? n : row_end) - k0 - jb, N = col_end - col_begin;
    if (M <= 0 || N <= 0) return;
    dim3 grid((M + kGM - 1) / kGM, (N + kGN - 1) / kGN);
    gemm_kernel<<<grid, 256, 0, s>>>(M, N, jb, A + (long long)k0 * lda + k0 + jb, A + (long long)col_begin * lda + k0,
                                     A + (long long)col_begin * lda + k0 + jb, lda);
    ++g_gpb_launches;
}

// Right-looking blocked LU with one-panel look-ahead: as soon as the columns of panel k+1 have received the update of
// panel k, panel k+1 is factored on a high-priority side stream (a 16-SM cluster) while the main stream finishes the
// trailing update of panel k on the rest of the matrix.
int factor_outer(int n, double* A, int lda, int* ipiv, int* info, double* B, int nrhs, int ldb, cudaStream_t s);
int outer_width(int n);

int factor_blocked(int n, double* A, int lda, int* ipiv, int* info, double* B, int nrhs, int ldb, cudaStream_t s) {
    // (the K = 256 update moves its operands with 16-byte copies when lda is even and A is 16-byte aligned: 7 % faster at
    // n = 35 000; the host mirror pads odd systems to an even leading dimension for that reason -- a stream-ordered
    // workspace allocated here instead cost 1.3 s per call at that size)
    if (outer_width(n) > 0) return factor_outer(n, A, lda, ipiv, info, B, nrhs, ldb, s);
    zero_info_kernel<<<1, 1, 0, s>>>(info);
    GPB_LAUNCH_CHECK();
    const LookAhead la = look_ahead();
    const bool ahead = la.ok;
    cudaStream_t ps = ahead ? la.panel_stream : s;
    const int rhs_blocks = (B != nullptr) ? (nrhs + 63) / 64 : 0;
    if (ahead) {                                     // the side stream starts after everything queued on s so far
        GPB_CHECK_CUDA(cudaEventRecord(la.ready, s));
        GPB_CHECK_CUDA(cudaStreamWaitEvent(ps, la.ready, 0));
    }
    int rc = launch_panel(n, 0, panel_width(n, 0), A, lda, ipiv, info, ps);
    if (rc) return rc;
    for (int k0 = 0; k0 < n;) {
        const int jb = panel_width(n, k0);
        const int k1 = k0 + jb;
        const int jb1 = (k1 < n) ? panel_width(n, k1) : 0;
        if (ahead) {                                 // panel k0 (side stream) -> main stream
            GPB_CHECK_CUDA(cudaEventRecord(la.panel_done, ps));
            GPB_CHECK_CUDA(cudaStreamWaitEvent(s, la.panel_done, 0));
        }
        const int nright = n - k1;
        const int mat_blocks = (nright + 63) / 64;
        if (mat_blocks + rhs_blocks > 0) {
            swap_trsm_kernel<<<mat_blocks + rhs_blocks, 256, 0, s>>>(n, k0, jb, A, lda, ipiv, mat_blocks, B, ldb, nrhs, k1, n, 1);
            GPB_LAUNCH_CHECK();
        }
        if (jb1 > 0) {
            // columns of the next panel first, then hand them to the side stream
            launch_gemm(n, k0, jb, k1, k1 + jb1, A, lda, s);
            if (ahead) {
                GPB_CHECK_CUDA(cudaEventRecord(la.ready, s));
                GPB_CHECK_CUDA(cudaStreamWaitEvent(ps, la.ready, 0));
            }
            if ((rc = launch_panel(n, k1, jb1, A, lda, ipiv, info, ps))) return rc;
            launch_gemm(n, k0, jb, k1 + jb1, n, A, lda, s);
            if (B != nullptr) {
                rhs_update_kernel<<<(n - k1 + 255) / 256, 256, 0, s>>>(n, k0, jb, A, lda, B, ldb, nrhs);
                GPB_LAUNCH_CHECK();
            }
        }
        k0 = k1;
    }
    if (ahead) {                                     // nothing of the side stream may outlive the call on s
        GPB_CHECK_CUDA(cudaEventRecord(la.panel_done, ps));
        GPB_CHECK_CUDA(cudaStreamWaitEvent(s, la.panel_done, 0));
    }
    GPB_CHECK_CUDA(cudaGetLastError());
    return GPB_OK;
}

// ---- outer-blocked variant for large n ----------------------------------------------------------------------
// With kNB-wide updates every panel reads and writes the whole trailing matrix: n^3/6 bytes in total, 1.1 s of HBM time at
// n = 35 000.  Above outer_min_n() the matrix is factored in outer blocks of kOuter columns: the block's own columns are
// factored panel by panel (interchanges LAPACK-style inside the block, so its L is in one row order), and the rest of
// the matrix sees ONE update per block with K = kOuter (gemm_big_kernel) -- 8x less traffic, compute bound.  The next
// block is factored on the side stream while the main stream finishes the big update (same look-ahead as above).
int g_outer_min_n = -1;
int outer_min_n() {
    if (g_outer_min_n < 0) {
        const char* e = getenv("GPB_LU_OUTER_MIN_N");
        g_outer_min_n = e ? atoi(e) : 10240;
    }
    return g_outer_min_n;
}
int g_outer_width = 0;           // 0: by n
int outer_width(int n) {
    if (n <= kSmallN || n < outer_min_n()) return 0;
    if (g_outer_width > 0) return g_outer_width;
    return kOuter;
}

int launch_big_gemm(int n, int K0, int W, int col_begin, int col_end, double* A, int lda, cudaStream_t s) {
    static bool attr_sets[GPB_MAX_DEVICES] = {false};
    bool& attr_set = attr_sets[gpb_current_device()];
    if (!attr_set) {
        GPB_CHECK_CUDA(cudaFuncSetAttribute(gemm_big_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigGemmSmem));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(gemm_big_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(gemm_big_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigGemmSmem));
        GPB_CHECK_CUDA(cudaFuncSetAttribute(gemm_big_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    const int K1 = K0 + W;
    const int M = n - K1, N = col_end - col_begin;
    if (M <= 0 || N <= 0) return GPB_OK;
    const int ntiles = ((M + kBM - 1) / kBM) * ((N + kBN - 1) / kBN);
    static const int tiles_per_cta = [] { const char* e = getenv("GPB_LU_GEMM_TILES"); const int v = e ? atoi(e) : kBTilesPerCta; return v > 0 ? v : 1; }();
    static const bool allow16 = getenv("GPB_LU_GEMM_NO16") == nullptr;
    const int grid = (ntiles + tiles_per_cta - 1) / tiles_per_cta;
    const double* Ap = A + (long long)K0 * lda + K1;
    const double* Bp = A + (long long)col_begin * lda + K0;
    double* Cp = A + (long long)col_begin * lda + K1;
    const bool aligned = allow16 && (lda % 2 == 0) && (reinterpret_cast<uintptr_t>(Ap) % 16 == 0) &&
                         (reinterpret_cast<uintptr_t>(Bp) % 16 == 0);
    if (aligned)
        gemm_big_kernel<true><<<grid, kBThreads, kBigGemmSmem, s>>>(M, N, W, Ap, Bp, Cp, lda, tiles_per_cta);
    else
        gemm_big_kernel<false><<<grid, kBThreads, kBigGemmSmem, s>>>(M, N, W, Ap, Bp, Cp, lda, tiles_per_cta);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

// columns [K0, K0+W): panel by panel, updates confined to the block's own columns
int factor_outer_panel(int n, int K0, int W, double* A, int lda, int* ipiv, int* info, cudaStream_t q) {
    for (int k0 = K0; k0 < K0 + W; k0 += kNB) {
        const int jb = min(kNB, K0 + W - k0);
        const int k1 = k0 + jb;
        int rc = launch_panel(n, k0, jb, A, lda, ipiv, info, q);
        if (rc) return rc;
        if (k0 > K0) {                               // earlier columns of the block follow the interchanges
            const int lb = (k0 - K0 + 255) / 256;
            laswp_kernel<<<lb, 256, 0, q>>>(k0, jb, A, lda, ipiv, K0, k0, lb, nullptr, 0, 0);
            GPB_LAUNCH_CHECK();
        }
        if (k1 < K0 + W) {
            const int nb = (K0 + W - k1 + 63) / 64;
            swap_trsm_kernel<<<nb, 256, 0, q>>>(n, k0, jb, A, lda, ipiv, nb, nullptr, 0, 0, k1, K0 + W, 1);
            GPB_LAUNCH_CHECK();
            launch_gemm(n, k0, jb, k1, K0 + W, A, lda, q);
        }
    }
    return GPB_OK;
}

int factor_outer(int n, double* A, int lda, int* ipiv, int* info, double* B, int nrhs, int ldb, cudaStream_t s) {
    zero_info_kernel<<<1, 1, 0, s>>>(info);
    GPB_LAUNCH_CHECK();
    const LookAhead la = look_ahead();
    const bool ahead = la.ok;
    cudaStream_t ps = ahead ? la.panel_stream : s;
    if (ahead) {
        GPB_CHECK_CUDA(cudaEventRecord(la.ready, s));
        GPB_CHECK_CUDA(cudaStreamWaitEvent(ps, la.ready, 0));
    }
    const int ow = outer_width(n);
    int rc = factor_outer_panel(n, 0, min(ow, n), A, lda, ipiv, info, ps);
    if (rc) return rc;
    for (int K0 = 0; K0 < n;) {
        const int W = min(ow, n - K0);
        const int K1 = K0 + W;
        const int W1 = (K1 < n) ? min(ow, n - K1) : 0;
        if (ahead) {
            GPB_CHECK_CUDA(cudaEventRecord(la.panel_done, ps));
            GPB_CHECK_CUDA(cudaStreamWaitEvent(s, la.panel_done, 0));
        }
        const int nright = n - K1;
        {   // all interchanges of the block on the columns right of it and on the right-hand sides
            const int mb = (nright + 255) / 256, rb = (B != nullptr) ? (nrhs + 255) / 256 : 0;
            if (mb + rb > 0) {
                laswp_kernel<<<mb + rb, 256, 0, s>>>(K0, W, A, lda, ipiv, K1, n, mb, B, ldb, nrhs);
                GPB_LAUNCH_CHECK();
            }
        }
        // U12 = L11^-1 A12 for the kOuter x kOuter unit-lower L11, panel by panel
        const int mat_blocks = (nright + 63) / 64;
        const int rhs_blocks = (B != nullptr) ? (nrhs + 63) / 64 : 0;
        for (int k0 = K0; k0 < K1; k0 += kNB) {
            const int jb = min(kNB, K1 - k0);
            const int k1 = k0 + jb;
            if (mat_blocks + rhs_blocks > 0) {
                swap_trsm_kernel<<<mat_blocks + rhs_blocks, 256, 0, s>>>(n, k0, jb, A, lda, ipiv, mat_blocks, B, ldb, nrhs, K1, n, 0);
                GPB_LAUNCH_CHECK();
            }
            if (k1 < K1 && nright > 0) launch_gemm(n, k0, jb, K1, n, A, lda, s, K1);
            if (B != nullptr && k1 < n) {
                rhs_update_kernel<<<(n - k1 + 255) / 256, 256, 0, s>>>(n, k0, jb, A, lda, B, ldb, nrhs);
                GPB_LAUNCH_CHECK();
            }
        }
        if (W1 > 0) {
            if ((rc = launch_big_gemm(n, K0, W, K1, K1 + W1, A, lda, s))) return rc;
            if (ahead) {
                GPB_CHECK_CUDA(cudaEventRecord(la.ready, s));
                GPB_CHECK_CUDA(cudaStreamWaitEvent(ps, la.ready, 0));
            }
            if ((rc = factor_outer_panel(n, K1, W1, A, lda, ipiv, info, ps))) return rc;
            if ((rc = launch_big_gemm(n, K0, W, K1 + W1, n, A, lda, s))) return rc;
        }
        K0 = K1;
    }
    if (ahead) {
        GPB_CHECK_CUDA(cudaEventRecord(la.panel_done, ps));
        GPB_CHECK_CUDA(cudaStreamWaitEvent(s, la.panel_done, 0));
    }
    GPB_CHECK_CUDA(cudaGetLastError());
    return GPB_OK;
}

// U x = y for every right-hand side (multi-CTA pipelined backward substitution)
int backward_blocked(int n, const double* LU, int lda, double* B, int nrhs, int ldb, cudaStream_t s) {
    const int nblk = (n + kNB - 1) / kNB;
    int* flags = nullptr;
    GPB_CHECK_CUDA(gpb_malloc_async((void**)&flags, sizeof(int) * nblk, s));
    for (int r = 0; r < nrhs; ++r) {
        GPB_CHECK_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * nblk, s));
        trsv_upper_kernel<<<nblk, 128, 0, s>>>(n, LU, lda, B + (long long)r * ldb, flags);
        GPB_LAUNCH_CHECK();
    }
    GPB_CHECK_CUDA(cudaFreeAsync(flags, s));
    return GPB_OK;
}

int small_path(int n, double* A, int lda, double* b, int nrhs, int ldb, int* ipiv, int* info, int do_solve, cudaStream_t s) {
    const size_t smem = (size_t)n * (n + 1) * sizeof(double);
    static bool attr_sets[GPB_MAX_DEVICES] = {false};
    bool& attr_set = attr_sets[gpb_current_device()];
    if (!attr_set) {
        GPB_CHECK_CUDA(cudaFuncSetAttribute(lu_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    lu_small_kernel<<<1, 256, smem, s>>>(n, A, lda, b, nrhs, ldb, ipiv, info, do_solve);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace

extern "C" int gpb_lu_set_outer_min_n(int min_n) {
    const int prev = outer_min_n();
    if (min_n >= 0) g_outer_min_n = min_n;
    return prev;
}

extern "C" int gpb_lu_set_outer_width(int width) {
    const int prev = g_outer_width;
    if (width == 0 || width == kOuter / 2 || width == kOuter) g_outer_width = width;
    return prev;
}

extern "C" int gpb_lu_factor(int n, double* A, int lda, int* ipiv, int* info, void* stream) {
    GPB_REQUIRE(n > 0 && A && ipiv && lda >= n, "bad arguments");
    if (n <= kSmallN) return small_path(n, A, lda, nullptr, 0, 0, ipiv, info, 0, (cudaStream_t)stream);
    GpbDeviceLock lock;
    return factor_blocked(n, A, lda, ipiv, info, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" int gpb_lu_apply(int n, const double* LU, int lda, const int* ipiv, double* b, int nrhs, int ldb, void* stream) {
    GPB_REQUIRE(n > 0 && LU && ipiv && b && lda >= n && ldb >= n && nrhs >= 1, "bad arguments");
    apply_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, LU, lda, ipiv, b, nrhs, ldb, outer_width(n));
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_lu_solve(int n, double* A, int lda, double* b, int nrhs, int ldb, int* ipiv, int* info, void* stream) {
    GPB_REQUIRE(n > 0 && A && b && ipiv && lda >= n && ldb >= n && nrhs >= 1, "bad arguments");
    if (n <= kSmallN) return small_path(n, A, lda, b, nrhs, ldb, ipiv, info, 1, (cudaStream_t)stream);
    // forward substitution rides along with the factorisation (b is one more block of columns), then U x = y
    GpbDeviceLock lock;
    int rc = factor_blocked(n, A, lda, ipiv, info, b, nrhs, ldb, (cudaStream_t)stream);
    if (rc) return rc;
    return backward_blocked(n, A, lda, b, nrhs, ldb, (cudaStream_t)stream);
}
