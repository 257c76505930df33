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
    int jb = kNB;
    if (cluster > 0) {
        const long long m = n - k0;
        const long long R = (m + cluster - 1) / cluster;
        while (jb > 8 && (unsigned long long)(R + 2) * jb * sizeof(double) > smem_cap) jb >>= 1;
    }
    return (n - k0 < jb) ? n - k0 : jb;
}

// Outer block at column k0 = two consecutive panels (w1 + w2 <= 64 columns).  Interchanges are LAPACK-style inside an
// outer block and LINPACK-style across outer blocks; the trailing update uses K = w1 + w2.  Systems that take the
// shared-memory path (n <= kSmallN) use single 32-column blocks.
__host__ __device__ inline void outer_widths_hd(int n, int k0, int cluster, unsigned long long smem_cap, int& w1, int& w2) {
    w1 = panel_width_hd(n, k0, cluster, smem_cap);
    w2 = 0;
    if (n > kSmallN && k0 + w1 < n) w2 = panel_width_hd(n, k0 + w1, cluster, smem_cap);
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
                     int R, int ldp) {
    extern __shared__ double P[];                  // [jb][ldp] : this CTA's rows of the panel, column-major
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

    for (int e = tid; e < nrows * jb; e += kPanelThreads) {
        const int c = e / nrows, i = e - c * nrows;
        P[c * ldp + i] = A[(long long)(k0 + c) * lda + r0 + i];
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
    for (int e = tid; e < nrows * jb; e += kPanelThreads) {
        const int c = e / nrows, i = e - c * nrows;
        A[(long long)(k0 + c) * lda + r0 + i] = P[c * ldp + i];
    }
    cluster.sync();                                // no CTA exits while a neighbour may still address its smem
}


// inv(L11) of the wb x wb unit-lower block at (k0, k0): one thread per column, forward substitution in shared
// memory.  U12 = L11^-1 A12 then becomes a small matrix product inside swap_trsm_kernel (the usual "TRSM through
// the inverse of the diagonal block"; |l_ij| <= 1 under partial pivoting, so the inverse is benign).
constexpr int kWB = 64;          // maximum outer block width
__global__ void __launch_bounds__(kWB) trinv_kernel(const double* __restrict__ A, int lda, int k0, int wb,
                                                    double* __restrict__ Linv) {
    // S[i][j], i > j : L(i, j);  S[j][i], i > j : X(i, j) = inv(L)(i, j)  (the inverse is unit lower too, so it fits
    // in the unused upper triangle, transposed: thread c owns row c of the upper part)
    __shared__ double S[kWB][kWB + 1];
    const int c = threadIdx.x;
    for (int e = c; e < wb * wb; e += kWB) {
        const int j = e / wb, i = e - j * wb;
        if (i > j) S[i][j] = A[(long long)(k0 + j) * lda + k0 + i];
    }
    __syncthreads();
    if (c < wb) {
        for (int i = c + 1; i < wb; ++i) {
            double acc = S[i][c];                       // k = c term: L(i, c) * X(c, c) = L(i, c)
            for (int k = c + 1; k < i; ++k) acc = fma(S[i][k], S[c][k], acc);
            S[c][i] = -acc;
        }
    }
    __syncthreads();
    for (int e = c; e < wb * wb; e += kWB) {
        const int j = e / wb, i = e - j * wb;
        Linv[j * kWB + i] = (i > j) ? S[j][i] : (i == j ? 1.0 : 0.0);     // column-major, ld = kWB
    }
}

// Apply the interchanges ipiv[k0 .. k0+wb) to the matrix columns [col_begin, col_end) and form
// U12 = inv(L11) * A12.  One CTA handles 64 columns; blocks beyond the matrix' own column blocks work on the
// right-hand sides B (n x nrhs, ldb): the forward substitution of gpb_lu_solve rides along with the factorisation.
//   1. the wb top rows of the 64 columns are staged in shared memory (coalesced);
//   2. the rows p >= k0 + wb touched by the interchanges are gathered in parallel (they are distinct unless
//      `dup` says otherwise, in which case the gather is serialised per column);
//   3. each column replays the interchange sequence in shared memory;
//   4. the displaced values go back to their rows, U12 = Linv * T is computed from shared memory (4 x 4 register
//      tiles) and written back.
__global__ void __launch_bounds__(256) swap_trsm_kernel(int n, int k0, int wb, int col_begin, int col_end,
                                                        double* __restrict__ A, int lda, const int* __restrict__ ipiv,
                                                        const double* __restrict__ Linv, int n_mat_blocks,
                                                        double* __restrict__ B, int ldb, int nrhs) {
    extern __shared__ double sm_st[];
    constexpr int LD = kWB + 1;
    double* Li = sm_st;                        // [k][i] : Linv(i, k) at Li[k * LD + i]
    double* T = sm_st + kWB * LD;              // [c][i] : top rows of column c
    double* G = sm_st + 2 * kWB * LD;          // [c][j] : value living at row ipiv[k0 + j] of column c
    __shared__ int piv[kWB];
    __shared__ int dup;
    const int tid = threadIdx.x;
    const bool on_rhs = (int)blockIdx.x >= n_mat_blocks;
    const int c0 = on_rhs ? ((int)blockIdx.x - n_mat_blocks) * 64 : col_begin + blockIdx.x * 64;
    const int ncol = on_rhs ? min(64, nrhs - c0) : min(64, col_end - c0);
    double* const M = on_rhs ? B : A;              // the columns this block transforms
    const int ldm = on_rhs ? ldb : lda;
    if (tid == 0) dup = 0;
    if (tid < wb) piv[tid] = ipiv[k0 + tid];
    for (int e = tid; e < wb * wb; e += 256) {
        const int k = e / wb, i = e - k * wb;
        Li[k * LD + i] = Linv[k * kWB + i];
    }
    for (int e = tid; e < ncol * wb; e += 256) {
        const int c = e / wb, i = e - c * wb;
        T[c * LD + i] = M[(long long)(c0 + c) * ldm + k0 + i];
    }
    __syncthreads();
    // a row below the block that is the target of two interchanges makes the gather order-dependent
    if (tid < wb && piv[tid] >= k0 + wb)
        for (int j2 = 0; j2 < tid; ++j2)
            if (piv[j2] == piv[tid]) dup = 1;
    __syncthreads();
    const bool serial = dup != 0;
    if (!serial) {
        for (int e = tid; e < ncol * wb; e += 256) {
            const int c = e / wb, j = e - c * wb;
            if (piv[j] >= k0 + wb) G[c * LD + j] = M[(long long)(c0 + c) * ldm + piv[j]];
        }
    }
    __syncthreads();
    if (tid < ncol) {
        const int c = tid;
        double* col = M + (long long)(c0 + c) * ldm;
        for (int j = 0; j < wb; ++j) {
            const int p = piv[j];
            if (p == k0 + j) continue;
            if (p < k0 + wb) {
                const double t = T[c * LD + j];
                T[c * LD + j] = T[c * LD + (p - k0)];
                T[c * LD + (p - k0)] = t;
            } else if (!serial) {
                const double t = T[c * LD + j];
                T[c * LD + j] = G[c * LD + j];
                G[c * LD + j] = t;
            } else {
                const double t = T[c * LD + j];
                T[c * LD + j] = col[p];
                col[p] = t;
            }
        }
    }
    __syncthreads();
    if (!serial) {
        for (int e = tid; e < ncol * wb; e += 256) {
            const int c = e / wb, j = e - c * wb;
            if (piv[j] >= k0 + wb && piv[j] != k0 + j) M[(long long)(c0 + c) * ldm + piv[j]] = G[c * LD + j];
        }
    }
    // U = Linv * T : thread (ti, tc) -> rows 4 ti .. 4 ti + 3, columns 4 tc .. 4 tc + 3
    const int ti = tid & 15, tc = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = 0.0;
    for (int k = 0; k < wb; ++k) {
        double lv[4], tv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) lv[a] = Li[k * LD + 4 * ti + a];
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2) tv[b2] = T[(4 * tc + b2) * LD + k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = fma(lv[a], tv[b2], acc[a][b2]);
    }
#pragma unroll
    for (int b2 = 0; b2 < 4; ++b2) {
        const int c = 4 * tc + b2;
        if (c < ncol) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int i = 4 * ti + a;
                if (i < wb) M[(long long)(c0 + c) * ldm + k0 + i] = acc[a][b2];
            }
        }
    }
}

// Apply the interchanges ipiv[ks .. ke) to the columns [c_begin, c_end) (second panel's swaps on the first
// panel's L columns: LAPACK-style inside an outer block).
__global__ void swap_cols_kernel(double* __restrict__ A, int lda, const int* __restrict__ ipiv, int ks, int ke, int c_begin, int c_end) {
    const int c = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_end) return;
    double* col = A + (long long)c * lda;
    for (int j = ks; j < ke; ++j) {
        const int p = ipiv[j];
        if (p != j) { const double t = col[j]; col[j] = col[p]; col[p] = t; }
    }
}

// Right-hand sides: B[k0+jb:, :] -= L21 * B[k0:k0+jb, :]   (the gemv twin of the trailing update)
__global__ void __launch_bounds__(256) rhs_update_kernel(int n, int k0, int jb, const double* __restrict__ A, int lda,
                                                         double* __restrict__ B, int ldb, int nrhs) {
    __shared__ double xs[kWB];
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

// ---- DMMA trailing update, 128 x 64 CTA tile, K in stages of 32 double-buffered with cp.async --------------
// acc = L21 * U12 is accumulated from zero in registers (4 x 4 m8n8 tiles per warp) while the next K stage streams
// into the other shared-memory buffer; C is read once at the end: C -= acc.
constexpr int kTM = 128, kTN = 64;
constexpr int kLdA2 = kTM + 8;          // 8 mod 16 doubles
constexpr int kLdB2 = kNB + 4;          // 4 mod 16 doubles
constexpr int kStageDoubles = kNB * kLdA2 + kTN * kLdB2;

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(gpb_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256) gemm128_kernel(int M, int N, int K, const double* __restrict__ Ap,
                                                      const double* __restrict__ Bp, double* __restrict__ C, int lda) {
    extern __shared__ double sm_g[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * kTM, n0 = blockIdx.y * kTN;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int r = lane >> 2, q = lane & 3;

    auto stage_in = [&](int st, int kh) {
        double* As = sm_g + st * kStageDoubles;
        double* Bs = As + kNB * kLdA2;
        const int kc = min(kNB, K - kh);
        for (int e = tid; e < kNB * kTM; e += 256) {
            const int k = e / kTM, m = e - k * kTM;
            if (k < kc && m0 + m < M) cp_async8(As + k * kLdA2 + m, Ap + (long long)(kh + k) * lda + m0 + m);
            else As[k * kLdA2 + m] = 0.0;
        }
        for (int e = tid; e < kTN * kNB; e += 256) {
            const int nn = e / kNB, k = e - nn * kNB;
            if (k < kc && n0 + nn < N) cp_async8(Bs + nn * kLdB2 + k, Bp + (long long)(n0 + nn) * lda + kh + k);
            else Bs[nn * kLdB2 + k] = 0.0;
        }
        cp_async_commit();
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int n_stages = (K + kNB - 1) / kNB;
    stage_in(0, 0);
    for (int st = 0; st < n_stages; ++st) {
        if (st + 1 < n_stages) {
            stage_in((st + 1) & 1, (st + 1) * kNB);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const double* As = sm_g + (st & 1) * kStageDoubles;
        const double* Bs = As + kNB * kLdA2;
#pragma unroll 2
        for (int ks = 0; ks < kNB; ks += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(ks + q) * kLdA2 + wm + 8 * i + r];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(wn + 8 * j + r) * kLdB2 + ks + q];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();               // the buffer may be refilled two stages later
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + wm + 8 * i + r;
            const int gn = n0 + wn + 8 * j + 2 * q;
            if (gm < M && gn < N) C[(long long)gn * lda + gm] -= acc[i][j][0];
            if (gm < M && gn + 1 < N) C[(long long)(gn + 1) * lda + gm] -= acc[i][j][1];
        }
}

// ---- triangular solves with the blocked factors (one CTA; panel-by-panel) --------------------------------
__global__ void __launch_bounds__(1024) apply_kernel(int n, const double* __restrict__ LU, int lda, const int* __restrict__ ipiv,
                                                      double* __restrict__ b, int nrhs, int ldb, int cluster,
                                                      unsigned long long smem_cap) {
    __shared__ double xs[kNB];
    __shared__ double D[kNB][kNB + 1];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int r = 0; r < nrhs; ++r) {
        double* x = b + (long long)r * ldb;
        // forward: P, L  (same outer-block schedule as the factorisation: all interchanges of an outer block first,
        // then its panels one after the other)
        int next_outer = 0;
        for (int k0 = 0, jb = 0; k0 < n; k0 += jb) {
            jb = panel_width_hd(n, k0, cluster, smem_cap);
            __syncthreads();
            if (k0 == next_outer) {
                int w1, w2;
                outer_widths_hd(n, k0, cluster, smem_cap, w1, w2);
                if (tid == 0) {
                    for (int j = k0; j < k0 + w1 + w2; ++j) {
                        const int p = ipiv[j];
                        if (p != j) { const double t = x[j]; x[j] = x[p]; x[p] = t; }
                    }
                }
                next_outer = k0 + w1 + w2;
            }
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
    static PanelConfig cfg;
    static bool done = false;
    if (done) return cfg;
    done = true;
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
bool panel_fits_cluster(int n, int k0, int jb) {
    const PanelConfig& cfg = panel_config();
    if (cfg.cluster == 0) return false;
    const long long m = n - k0;
    const long long R = (m + cfg.cluster - 1) / cfg.cluster;
    return (size_t)(R + 2) * jb * sizeof(double) <= cfg.smem_cap;
}

int launch_panel(int n, int k0, int jb, double* A, int lda, int* ipiv, int* info, cudaStream_t s) {
    if (!panel_fits_cluster(n, k0, jb)) {
        panel_kernel<<<1, 1024, 0, s>>>(n, k0, jb, A, lda, ipiv, info);
        GPB_LAUNCH_CHECK();
        return GPB_OK;
    }
    const PanelConfig& cfg = panel_config();
    const int m = n - k0;
    int R = (m + cfg.cluster - 1) / cfg.cluster;
    if (R < 1) R = 1;
    const int ldp = (R + 1) & ~1;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(cfg.cluster);
    lc.blockDim = dim3(kPanelThreads);
    lc.dynamicSmemBytes = (size_t)ldp * jb * sizeof(double);
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cfg.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    GPB_CHECK_CUDA(cudaLaunchKernelEx(&lc, panel_cluster_kernel, n, k0, jb, A, lda, ipiv, info, R, ldp));
    ++g_gpb_launches;
    return GPB_OK;
}


int launch_swap_trsm(int n, int k0, int wb, int col_begin, int col_end, double* A, int lda, const int* ipiv, double* Linv,
                     double* B, int ldb, int nrhs, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CHECK_CUDA(cudaFuncSetAttribute(swap_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        attr_set = true;
    }
    const int mat_blocks = (col_end > col_begin) ? (col_end - col_begin + 63) / 64 : 0;
    const int rhs_blocks = (B != nullptr) ? (nrhs + 63) / 64 : 0;
    if (mat_blocks + rhs_blocks == 0) return GPB_OK;
    trinv_kernel<<<1, kWB, 0, s>>>(A, lda, k0, wb, Linv);
    GPB_LAUNCH_CHECK();
    const size_t smem = (size_t)3 * kWB * (kWB + 1) * sizeof(double);
    swap_trsm_kernel<<<mat_blocks + rhs_blocks, 256, smem, s>>>(n, k0, wb, col_begin, col_end, A, lda, ipiv, Linv, mat_blocks,
                                                                B, ldb, nrhs);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int launch_gemm(int n, int k0, int K, int col_begin, int col_end, double* A, int lda, cudaStream_t s) {
    // C[k0+K:, col_begin:col_end] -= A[k0+K:, k0:k0+K] * A[k0:k0+K, col_begin:col_end]
    const int M = n - k0 - K, N = col_end - col_begin;
    if (M <= 0 || N <= 0) return GPB_OK;
    static bool attr_set = false;
    const size_t smem = (size_t)2 * kStageDoubles * sizeof(double);
    if (!attr_set) {
        GPB_CHECK_CUDA(cudaFuncSetAttribute(gemm128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid((M + kTM - 1) / kTM, (N + kTN - 1) / kTN);
    gemm128_kernel<<<grid, 256, smem, s>>>(M, N, K, A + (long long)k0 * lda + k0 + K, A + (long long)col_begin * lda + k0,
                                           A + (long long)col_begin * lda + k0 + K, lda);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

int factor_blocked(int n, double* A, int lda, int* ipiv, int* info, double* B, int nrhs, int ldb, cudaStream_t s) {
    zero_info_kernel<<<1, 1, 0, s>>>(info);
    GPB_LAUNCH_CHECK();
    const PanelConfig& cfg = panel_config();
    int rc;
    double* Linv = nullptr;                          // inverse of the current diagonal block (kWB x kWB)
    GPB_CHECK_CUDA(cudaMallocAsync((void**)&Linv, sizeof(double) * kWB * kWB, s));
    for (int k0 = 0; k0 < n;) {
        int w1, w2;
        outer_widths_hd(n, k0, cfg.cluster, (unsigned long long)cfg.smem_cap, w1, w2);
        const int wb = w1 + w2;
        // ---- first panel
        if ((rc = launch_panel(n, k0, w1, A, lda, ipiv, info, s))) return rc;
        if (w2 > 0) {
            // ---- bring the second panel's columns up to date (K = w1), factor it, complete the interchanges of
            //      the first panel's L columns (LAPACK-style inside the outer block)
            if ((rc = launch_swap_trsm(n, k0, w1, k0 + w1, k0 + wb, A, lda, ipiv, Linv, nullptr, 0, 0, s))) return rc;
            if ((rc = launch_gemm(n, k0, w1, k0 + w1, k0 + wb, A, lda, s))) return rc;
            if ((rc = launch_panel(n, k0 + w1, w2, A, lda, ipiv, info, s))) return rc;
            swap_cols_kernel<<<(w1 + 63) / 64, 64, 0, s>>>(A, lda, ipiv, k0 + w1, k0 + wb, k0, k0 + w1);
            GPB_LAUNCH_CHECK();
        }
        // ---- everything right of the outer block (and the right-hand sides): interchanges, U12, trailing update (K = wb)
        if ((rc = launch_swap_trsm(n, k0, wb, k0 + wb, n, A, lda, ipiv, Linv, B, ldb, nrhs, s))) return rc;
        if ((rc = launch_gemm(n, k0, wb, k0 + wb, n, A, lda, s))) return rc;
        if (B != nullptr && n - k0 - wb > 0) {
            rhs_update_kernel<<<(n - k0 - wb + 255) / 256, 256, 0, s>>>(n, k0, wb, A, lda, B, ldb, nrhs);
            GPB_LAUNCH_CHECK();
        }
        k0 += wb;
    }
    GPB_CHECK_CUDA(cudaFreeAsync(Linv, s));
    return GPB_OK;
}

// U x = y for every right-hand side (multi-CTA pipelined backward substitution)
int backward_blocked(int n, const double* LU, int lda, double* B, int nrhs, int ldb, cudaStream_t s) {
    const int nblk = (n + kNB - 1) / kNB;
    int* flags = nullptr;
    GPB_CHECK_CUDA(cudaMallocAsync((void**)&flags, sizeof(int) * nblk, s));
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
    static bool attr_set = false;
    if (!attr_set) {
        GPB_CHECK_CUDA(cudaFuncSetAttribute(lu_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    lu_small_kernel<<<1, 256, smem, s>>>(n, A, lda, b, nrhs, ldb, ipiv, info, do_solve);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

}  // namespace

extern "C" int gpb_lu_factor(int n, double* A, int lda, int* ipiv, int* info, void* stream) {
    GPB_REQUIRE(n > 0 && A && ipiv && lda >= n, "bad arguments");
    if (n <= kSmallN) return small_path(n, A, lda, nullptr, 0, 0, ipiv, info, 0, (cudaStream_t)stream);
    return factor_blocked(n, A, lda, ipiv, info, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" int gpb_lu_apply(int n, const double* LU, int lda, const int* ipiv, double* b, int nrhs, int ldb, void* stream) {
    GPB_REQUIRE(n > 0 && LU && ipiv && b && lda >= n && ldb >= n && nrhs >= 1, "bad arguments");
    const PanelConfig& cfg = panel_config();
    apply_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, LU, lda, ipiv, b, nrhs, ldb, cfg.cluster,
                                                       (unsigned long long)cfg.smem_cap);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_lu_solve(int n, double* A, int lda, double* b, int nrhs, int ldb, int* ipiv, int* info, void* stream) {
    GPB_REQUIRE(n > 0 && A && b && ipiv && lda >= n && ldb >= n && nrhs >= 1, "bad arguments");
    if (n <= kSmallN) return small_path(n, A, lda, b, nrhs, ldb, ipiv, info, 1, (cudaStream_t)stream);
    // forward substitution rides along with the factorisation (b is one more block of columns), then U x = y
    int rc = factor_blocked(n, A, lda, ipiv, info, b, nrhs, ldb, (cudaStream_t)stream);
    if (rc) return rc;
    return backward_blocked(n, A, lda, b, nrhs, ldb, (cudaStream_t)stream);
}
