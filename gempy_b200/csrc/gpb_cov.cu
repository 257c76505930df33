// (1) Covariance assembly of the saddle-point co-kriging system.
//
// Engine stage replaced: "kernel_constructor" (yield_covariance) -- SURVEY.md 8a2 row (1); reference call
// site gempy/API/compute_API.py:68-73.  The reference builds ~20 n x n temporaries with numpy broadcasting;
// here every 32 x 32 tile is produced in one pass: the descriptors (class, axis, two coordinate triples,
// nugget) of the tile's 32 rows and 32 columns are staged in shared memory with coalesced FP64 loads, every
// thread computes 4 entries, and the tile is written once with coalesced column-major stores.
// Roofline: HBM write, 8 B per entry (n^2 * 8 B per launch).
//
// Entry formulas: oracle/gempy_oracle.py header ("Conventions"); constants pinned by the reference's
// approved scalar-field vectors.
#include "gpb_common.cuh"
#include <cstdlib>

namespace {

constexpr int kTile = 32;
enum RowClass : int { kG = 0, kI = 1, kU = 2, kF = 3 };

struct RowDesc {
    double a[3];     // G: orientation position; I: rest point; U/F: unused
    double b[3];     // I: reference point
    double nug;      // diagonal nugget (already the row nugget, NOT yet multiplied by c_o)
    int cls;         // RowClass
    int idx;         // G: axis (0..2); I: increment index; U: drift term; F: fault column
};

__device__ __forceinline__ void load_desc(const gpb_stack& st, int i, RowDesc& d) {
    const int n_g = 3 * st.n_ori;
    if (i < n_g) {
        const int ax = i / st.n_ori, o = i - ax * st.n_ori;
        d.cls = kG; d.idx = ax;
        d.a[0] = st.ori_pos[o]; d.a[1] = st.ori_pos[st.n_ori + o]; d.a[2] = st.ori_pos[2 * st.n_ori + o];
        d.b[0] = d.b[1] = d.b[2] = 0.0;
        d.nug = st.ori_nugget[o];
    } else if (i < n_g + st.n_rest) {
        const int r = i - n_g;
        d.cls = kI; d.idx = r;
        d.a[0] = st.rest[r]; d.a[1] = st.rest[st.n_rest + r]; d.a[2] = st.rest[2 * st.n_rest + r];
        d.b[0] = st.ref[r]; d.b[1] = st.ref[st.n_rest + r]; d.b[2] = st.ref[2 * st.n_rest + r];
        d.nug = st.sp_nugget[r];
    } else if (i < n_g + st.n_rest + st.n_drift) {
        d.cls = kU; d.idx = i - n_g - st.n_rest;
        d.a[0] = d.a[1] = d.a[2] = d.b[0] = d.b[1] = d.b[2] = 0.0; d.nug = 0.0;
    } else {
        d.cls = kF; d.idx = i - n_g - st.n_rest - st.n_drift;
        d.a[0] = d.a[1] = d.a[2] = d.b[0] = d.b[1] = d.b[2] = 0.0; d.nug = 0.0;
    }
}

// C(r), C'(r)/r, C''(r) for c_o = 1
template <int KERNEL>
__device__ __forceinline__ void cov_terms(double r, double a, double& C, double& kp, double& ka) {
    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {
        const double t = r / a, t2 = t * t;
        C = 1.0 - 7.0 * t2 + 8.75 * t2 * t - 3.5 * t2 * t2 * t + 0.75 * t2 * t2 * t2 * t;
        kp = (-14.0 + 26.25 * t - 17.5 * t2 * t + 5.25 * t2 * t2 * t) / (a * a);
        ka = 7.0 * (9.0 * t2 * t2 * t - 20.0 * t2 * t + 15.0 * t - 4.0) / (2.0 * a * a);
    } else if constexpr (KERNEL == GPB_KERNEL_EXPONENTIAL) {
        const double e = exp(-(r * r) / (2.0 * a * a));
        C = e; kp = -e / (a * a); ka = e * (r * r / (a * a * a * a) - 1.0 / (a * a));
    } else {
        const double s = 2.23606797749978969641 * r / a, e = exp(-s);
        C = (1.0 + s + s * s / 3.0) * e;
        kp = -(5.0 / (3.0 * a * a)) * (1.0 + s) * e;
        ka = -(5.0 / (3.0 * a * a)) * (1.0 + s - s * s) * e;
    }
}

__device__ __forceinline__ double dist3(const double* p, const double* q, double* h) {
    h[0] = p[0] - q[0]; h[1] = p[1] - q[1]; h[2] = p[2] - q[2];
    return sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2] + GPB_DIST_EPS);
}

__device__ __forceinline__ double drift_f(const double* x, int k) {
    switch (k) {
        case 0: return x[0]; case 1: return x[1]; case 2: return x[2];
        case 3: return x[0] * x[0]; case 4: return x[1] * x[1]; case 5: return x[2] * x[2];
        case 6: return x[0] * x[1]; case 7: return x[0] * x[2]; default: return x[1] * x[2];
    }
}
__device__ __forceinline__ double drift_df(const double* x, int k, int ax) {
    if (k < 3) return k == ax ? 1.0 : 0.0;
    if (k < 6) return (k - 3) == ax ? 2.0 * x[ax] : 0.0;
    const int p = (k == 8) ? 1 : 0, q = (k == 6) ? 1 : 2;      // xy:(0,1) xz:(0,2) yz:(1,2)
    if (ax == p) return x[q];
    if (ax == q) return x[p];
    return 0.0;
}

template <int KERNEL>
__device__ double entry(const gpb_stack& st, const RowDesc& ri, const RowDesc& rj, bool diag) {
    const RowDesc& p = (ri.cls <= rj.cls) ? ri : rj;     // lower class first
    const RowDesc& q = (ri.cls <= rj.cls) ? rj : ri;
    const double a = st.range, c_o = st.c_o;
    double h[3], C, kp, ka;
    if (p.cls == kG && q.cls == kG) {
        const double r = dist3(p.a, q.a, h);
        cov_terms<KERNEL>(r, a, C, kp, ka);
        double v = h[p.idx] * h[q.idx] * ((kp - ka) / (r * r + GPB_REG_EPS));
        if (p.idx == q.idx) v -= kp;
        v *= c_o;
        if (diag) v += c_o * p.nug;
        return v;
    }
    if (p.cls == kG && q.cls == kI) {
        const double r1 = dist3(p.a, q.a, h);
        cov_terms<KERNEL>(r1, a, C, kp, ka);
        double v = h[p.idx] * kp;
        const double r0 = dist3(p.a, q.b, h);
        cov_terms<KERNEL>(r0, a, C, kp, ka);
        v -= h[p.idx] * kp;
        return c_o * st.gi_res * v;
    }
    if (p.cls == kI && q.cls == kI) {
        double C11, C10, C01, C00;
        cov_terms<KERNEL>(dist3(p.a, q.a, h), a, C11, kp, ka);
        cov_terms<KERNEL>(dist3(p.a, q.b, h), a, C10, kp, ka);
        cov_terms<KERNEL>(dist3(p.b, q.a, h), a, C01, kp, ka);
        cov_terms<KERNEL>(dist3(p.b, q.b, h), a, C00, kp, ka);
        double v = c_o * st.i_res * ((C11 + C00) - (C10 + C01));      // symmetric under (i, j) swap
        if (diag) v += c_o * p.nug;
        return v;
    }
    if (p.cls == kG && q.cls == kU) return drift_df(p.a, q.idx, p.idx);
    if (p.cls == kI && q.cls == kU) return st.gi_res * (drift_f(p.a, q.idx) - drift_f(p.b, q.idx));
    if (p.cls == kI && q.cls == kF)
        return st.fault_rest[(long long)q.idx * st.n_rest + p.idx] - st.fault_ref[(long long)q.idx * st.n_rest + p.idx];
    return 0.0;     // G-F, U-U, U-F, F-F
}

template <int KERNEL>
__global__ void __launch_bounds__(256) cov_kernel(const gpb_stack st, int n, double* __restrict__ A, int lda,
                                                  double* __restrict__ b) {
    __shared__ RowDesc rows[kTile];
    __shared__ RowDesc cols[kTile];
    const int i0 = blockIdx.x * kTile, j0 = blockIdx.y * kTile;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid < kTile) {
        if (i0 + tid < n) load_desc(st, i0 + tid, rows[tid]);
    } else if (tid < 2 * kTile) {
        if (j0 + tid - kTile < n) load_desc(st, j0 + tid - kTile, cols[tid - kTile]);
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i < n) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jl = threadIdx.y + 8 * q;
            const int j = j0 + jl;
            if (j < n) A[(long long)j * lda + i] = entry<KERNEL>(st, rows[threadIdx.x], cols[jl], i == j);
        }
    }
    // right-hand side: b = [G_x; G_y; G_z; 0]
    if (blockIdx.y == 0 && threadIdx.y == 0 && i < n && b != nullptr)
        b[i] = (i < 3 * st.n_ori) ? st.ori_grad[i] : 0.0;      // ori_grad is [3][n_ori] = exactly this order
}


// =====================================================================================================================
// Blocked assembly for larger systems: one kernel per block class, every distance computed once.
//   cov_ii_kernel  increments x increments: 4 distances -> 1 entry (C only); lower tiles, mirrored through shared memory
//   cov_ig_kernel  increments x orientations: 2 distances -> 3 entries (one per gradient axis), + mirrored upper block
//   cov_gg_kernel  orientations x orientations: 1 distance -> 9 entries (3 x 3 axis blocks)
//   cov_du_kernel  universal-drift and fault-drift rows / columns, zero corner, right-hand side
// All coordinates are divided by the range once (no division per entry; __dmul_rn so that the product is rounded before it
// meets a subtraction: an FMA-contracted normalisation would make the (i, j) and (j, i) entries, which different threads
// compute, differ in the last bit), sqrt is the MUFU-seeded gpb_fast_sqrt.
// Roofline: HBM write, 8 B per entry; the arithmetic (about 64 FP64 operations per increment-increment entry) stays under
// it once nothing is recomputed.
// =====================================================================================================================
constexpr int kBT = 32;

// range-normalised kernel terms: u = r^2 / a^2 (with the distance epsilon), t = sqrt(u)
//   C(r), kp_n = a^2 C'(r)/r, dd_n = a^2 (C'(r)/r - C''(r))
template <int KERNEL>
__device__ __forceinline__ double cov_c_n(double u, double t) {
    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {
        return fma(t * u, fma(u, fma(0.75, u, -3.5), 8.75), fma(-7.0, u, 1.0));
    } else if constexpr (KERNEL == GPB_KERNEL_EXPONENTIAL) {
        return exp(-0.5 * u);
    } else {
        const double s = 2.23606797749978969641 * t;
        return fma(s, fma(s, 1.0 / 3.0, 1.0), 1.0) * exp(-s);
    }
}
template <int KERNEL>
__device__ __forceinline__ double cov_kp_n(double u, double t) {
    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {
        return fma(t, fma(u, fma(5.25, u, -17.5), 26.25), -14.0);
    } else if constexpr (KERNEL == GPB_KERNEL_EXPONENTIAL) {
        return -exp(-0.5 * u);
    } else {
        const double s = 2.23606797749978969641 * t;
        return (-5.0 / 3.0) * (1.0 + s) * exp(-s);
    }
}
template <int KERNEL>
__device__ __forceinline__ void cov_kp_dd_n(double u, double t, double& kp, double& dd) {
    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {
        kp = fma(t, fma(u, fma(5.25, u, -17.5), 26.25), -14.0);
        const double om = 1.0 - u;
        dd = -26.25 * t * om * om;                         // a^2 (C'/r - C'') = -(105/4) t (1 - u)^2
    } else if constexpr (KERNEL == GPB_KERNEL_EXPONENTIAL) {
        const double e = exp(-0.5 * u);
        kp = -e;
        dd = -e * u;
    } else {
        const double s = 2.23606797749978969641 * t;
        const double e = exp(-s);
        kp = (-5.0 / 3.0) * (1.0 + s) * e;
        dd = (-5.0 / 3.0) * e * s * s;
    }
}

__device__ __forceinline__ double u_of(double ax, double ay, double az, double bx, double by, double bz, double eps_u) {
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    return fma(dz, dz, fma(dy, dy, fma(dx, dx, eps_u)));
}

struct CovGeom {
    int n_ori, n_rest, n_g;        // n_g = 3 n_ori
    double inv_a, eps_u, eps_reg, c_o, c_i, c_gi, c_gg;     // c_i = c_o i_res, c_gi = c_o gi_res / a, c_gg = c_o / a^2
};

// increments x increments.  grid.x enumerates the lower tiles (ti >= tj) of the n_rest x n_rest block.
template <int KERNEL>
__global__ void __launch_bounds__(256) cov_ii_kernel(const gpb_stack st, const CovGeom g, double* __restrict__ A, int lda, int lower_only) {
    __shared__ double cr[6][kBT];                  // column points: rest xyz, ref xyz (range-normalised)
    __shared__ double S[kBT][kBT + 1];
    // triangular decode: t = ti (ti + 1) / 2 + tj
    const long long t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)ti * (ti + 1) / 2 > t) --ti;
    while ((long long)(ti + 1) * (ti + 2) / 2 <= t) ++ti;
    const int tj = (int)(t - (long long)ti * (ti + 1) / 2);
    const int i0 = ti * kBT, j0 = tj * kBT;
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 32 + lane;
    const int nr = g.n_rest;
    if (tid < 6 * kBT) {
        const int c = tid / kBT, jl = tid - c * kBT;
        const int j = min(j0 + jl, nr - 1);
        const double* src = (c < 3) ? st.rest + (long long)c * nr : st.ref + (long long)(c - 3) * nr;
        cr[c][jl] = __dmul_rn(src[j], g.inv_a);
    }
    const int i = i0 + lane;
    const int ic = min(i, nr - 1);
    const double rx = __dmul_rn(st.rest[ic], g.inv_a), ry = __dmul_rn(st.rest[nr + ic], g.inv_a), rz = __dmul_rn(st.rest[2LL * nr + ic], g.inv_a);
    const double fx = __dmul_rn(st.ref[ic], g.inv_a), fy = __dmul_rn(st.ref[nr + ic], g.inv_a), fz = __dmul_rn(st.ref[2LL * nr + ic], g.inv_a);
    const double nug = g.c_o * st.sp_nugget[ic];
    __syncthreads();
    double* const Ab = A + (long long)g.n_g * lda + g.n_g;          // the block's (0, 0)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jl = ty + 8 * q;
        const int j = j0 + jl;
        const double u11 = u_of(rx, ry, rz, cr[0][jl], cr[1][jl], cr[2][jl], g.eps_u);
        const double u10 = u_of(rx, ry, rz, cr[3][jl], cr[4][jl], cr[5][jl], g.eps_u);
        const double u01 = u_of(fx, fy, fz, cr[0][jl], cr[1][jl], cr[2][jl], g.eps_u);
        const double u00 = u_of(fx, fy, fz, cr[3][jl], cr[4][jl], cr[5][jl], g.eps_u);
        const double C11 = cov_c_n<KERNEL>(u11, gpb_fast_sqrt(u11));
        const double C10 = cov_c_n<KERNEL>(u10, gpb_fast_sqrt(u10));
        const double C01 = cov_c_n<KERNEL>(u01, gpb_fast_sqrt(u01));
        const double C00 = cov_c_n<KERNEL>(u00, gpb_fast_sqrt(u00));
        double v = g.c_i * ((C11 + C00) - (C10 + C01));              // symmetric under (i, j) swap, bit for bit
        if (i == j) v += nug;
        S[jl][lane] = v;
        if (i < nr && j < nr && (ti > tj || i >= j || !lower_only)) Ab[(long long)j * lda + i] = v;
    }
    if (ti == tj || lower_only) return;
    __syncthreads();
    // mirrored tile: A[j, i] = A[i, j], written with the tile's column index along the lanes
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int il = ty + 8 * q;                 // row of the original tile = column of the mirrored one
        const int ii = i0 + il, jj = j0 + lane;
        if (ii < nr && jj < nr) Ab[(long long)ii * lda + jj] = S[lane][il];
    }
}

// increments (rows, lanes) x orientations (columns): entry (I_i, G_(o, a)) = c_gi ((x_o - rest_i)_a kp(o, rest_i) - (x_o - ref_i)_a kp(o, ref_i))
template <int KERNEL>
__global__ void __launch_bounds__(256) cov_ig_kernel(const gpb_stack st, const CovGeom g, double* __restrict__ A, int lda, int lower_only) {
    __shared__ double co[3][kBT];                  // column points: orientation positions (range-normalised)
    __shared__ double S[3][kBT][kBT + 1];
    const int i0 = blockIdx.x * kBT, o0 = blockIdx.y * kBT;
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 32 + lane;
    const int nr = g.n_rest, no = g.n_ori;
    if (tid < 3 * kBT) {
        const int c = tid / kBT, ol = tid - c * kBT;
        co[c][ol] = __dmul_rn(st.ori_pos[(long long)c * no + min(o0 + ol, no - 1)], g.inv_a);
    }
    const int i = i0 + lane;
    const int ic = min(i, nr - 1);
    const double rx = __dmul_rn(st.rest[ic], g.inv_a), ry = __dmul_rn(st.rest[nr + ic], g.inv_a), rz = __dmul_rn(st.rest[2LL * nr + ic], g.inv_a);
    const double fx = __dmul_rn(st.ref[ic], g.inv_a), fy = __dmul_rn(st.ref[nr + ic], g.inv_a), fz = __dmul_rn(st.ref[2LL * nr + ic], g.inv_a);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ol = ty + 8 * q;
        const int o = o0 + ol;
        const double ox = co[0][ol], oy = co[1][ol], oz = co[2][ol];
        const double u1 = u_of(ox, oy, oz, rx, ry, rz, g.eps_u), u0 = u_of(ox, oy, oz, fx, fy, fz, g.eps_u);
        const double k1 = cov_kp_n<KERNEL>(u1, gpb_fast_sqrt(u1)), k0 = cov_kp_n<KERNEL>(u0, gpb_fast_sqrt(u0));
        const double vx = g.c_gi * ((ox - rx) * k1 - (ox - fx) * k0);
        const double vy = g.c_gi * ((oy - ry) * k1 - (oy - fy) * k0);
        const double vz = g.c_gi * ((oz - rz) * k1 - (oz - fz) * k0);
        S[0][ol][lane] = vx; S[1][ol][lane] = vy; S[2][ol][lane] = vz;
        if (i < nr && o < no) {                    // lower block: row n_g + i, column a n_ori + o
            double* p = A + (long long)o * lda + g.n_g + i;
            p[0] = vx;
            p[(long long)no * lda] = vy;
            p[2LL * no * lda] = vz;
        }
    }
    if (lower_only) return;
    __syncthreads();
    // mirrored (upper) block: row a n_ori + o, column n_g + i, written with o along the lanes
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int il = ty + 8 * q;
        const int ii = i0 + il, oo = o0 + lane;
        if (ii < nr && oo < no) {
            double* p = A + (long long)(g.n_g + ii) * lda + oo;
            p[0] = S[0][lane][il];
            p[no] = S[1][lane][il];
            p[2 * no] = S[2][lane][il];
        }
    }
}

// orientations x orientations: one distance per pair, nine entries (full block: it is 3 n_ori squared, small next to the rest)
template <int KERNEL>
__global__ void __launch_bounds__(256) cov_gg_kernel(const gpb_stack st, const CovGeom g, double* __restrict__ A, int lda) {
    __shared__ double co[3][kBT];
    const int o0 = blockIdx.x * kBT, p0 = blockIdx.y * kBT;
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 32 + lane;
    const int no = g.n_ori;
    if (tid < 3 * kBT) {
        const int c = tid / kBT, pl = tid - c * kBT;
        co[c][pl] = __dmul_rn(st.ori_pos[(long long)c * no + min(p0 + pl, no - 1)], g.inv_a);
    }
    const int o = o0 + lane;
    const int oc = min(o, no - 1);
    const double ox = __dmul_rn(st.ori_pos[oc], g.inv_a), oy = __dmul_rn(st.ori_pos[no + oc], g.inv_a), oz = __dmul_rn(st.ori_pos[2LL * no + oc], g.inv_a);
    const double nug = g.c_o * st.ori_nugget[oc];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int pl = ty + 8 * q;
        const int p = p0 + pl;
        const double hx = ox - co[0][pl], hy = oy - co[1][pl], hz = oz - co[2][pl];
        const double u = fma(hz, hz, fma(hy, hy, fma(hx, hx, g.eps_u)));
        double kp, dd;
        cov_kp_dd_n<KERNEL>(u, gpb_fast_sqrt(u), kp, dd);
        const double T = dd * gpb_fast_rcp(u + g.eps_reg);          // (C'/r - C'') / (r^2 + 1e-5), range-normalised
        const double kpc = g.c_gg * kp;
        const double Tc = g.c_gg * T;
        const double d0 = (o == p) ? nug : 0.0;
        if (o < no && p < no) {
            const double h[3] = {hx, hy, hz};
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    double v = h[a] * h[b] * Tc;
                    if (a == b) v = v - kpc + d0;
                    A[(long long)(b * no + p) * lda + a * no + o] = v;
                }
        }
    }
}

// drift and fault rows / columns + zero corner + right-hand side
__global__ void cov_du_kernel(const gpb_stack st, int n, double* __restrict__ A, int lda, double* __restrict__ b, int lower_only) {
    const int n_g = 3 * st.n_ori, nk = n_g + st.n_rest, nd = st.n_drift + st.n_faults;
    const long long total = (long long)n * nd;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(e / n), r = (int)(e - (long long)k * n);      // column nk + k, row r
        double v = 0.0;
        if (r < n_g) {
            if (k < st.n_drift) {
                const int ax = r / st.n_ori, o = r - ax * st.n_ori;
                const double x[3] = {st.ori_pos[o], st.ori_pos[st.n_ori + o], st.ori_pos[2LL * st.n_ori + o]};
                v = drift_df(x, k, ax);
            }
        } else if (r < nk) {
            const int i = r - n_g;
            if (k < st.n_drift) {
                const double xr[3] = {st.rest[i], st.rest[st.n_rest + i], st.rest[2LL * st.n_rest + i]};
                const double xf[3] = {st.ref[i], st.ref[st.n_rest + i], st.ref[2LL * st.n_rest + i]};
                v = st.gi_res * (drift_f(xr, k) - drift_f(xf, k));
            } else {
                const int f = k - st.n_drift;
                v = st.fault_rest[(long long)f * st.n_rest + i] - st.fault_ref[(long long)f * st.n_rest + i];
            }
        }
        if (!lower_only || r >= nk) A[(long long)(nk + k) * lda + r] = v;        // column nk + k (upper part + corner)
        if (r < nk) A[(long long)r * lda + nk + k] = v;                           // row nk + k (lower part)
    }
    if (b != nullptr)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            b[i] = (i < n_g) ? st.ori_grad[i] : 0.0;
}

template <int KERNEL>
int launch_blocked(const gpb_stack* st, int n, double* A, int lda, double* b, int lower_only, cudaStream_t s) {
    CovGeom g;
    g.n_ori = st->n_ori; g.n_rest = st->n_rest; g.n_g = 3 * st->n_ori;
    const double a = st->range;
    g.inv_a = 1.0 / a;
    g.eps_u = GPB_DIST_EPS / (a * a);
    g.eps_reg = GPB_REG_EPS / (a * a);
    g.c_o = st->c_o;
    g.c_i = st->c_o * st->i_res;
    g.c_gi = st->c_o * st->gi_res / a;
    g.c_gg = st->c_o / (a * a);
    const dim3 block(32, 8);
    const int To = (st->n_ori + kBT - 1) / kBT, Tr = (st->n_rest + kBT - 1) / kBT;
    if (Tr > 0) {
        const long long tiles = (long long)Tr * (Tr + 1) / 2;
        cov_ii_kernel<KERNEL><<<(unsigned)tiles, block, 0, s>>>(*st, g, A, lda, lower_only);
        GPB_LAUNCH_CHECK();
    }
    if (Tr > 0 && To > 0) {
        cov_ig_kernel<KERNEL><<<dim3(Tr, To), block, 0, s>>>(*st, g, A, lda, lower_only);
        GPB_LAUNCH_CHECK();
    }
    if (To > 0) {
        cov_gg_kernel<KERNEL><<<dim3(To, To), block, 0, s>>>(*st, g, A, lda);
        GPB_LAUNCH_CHECK();
    }
    {
        const long long work = (long long)n * max(1, st->n_drift + st->n_faults);
        long long nb = (work + 255) / 256;
        const int blocks = (int)(nb > 4096 ? 4096 : (nb < 1 ? 1 : nb));
        cov_du_kernel<<<blocks, 256, 0, s>>>(*st, n, A, lda, b, lower_only);
        GPB_LAUNCH_CHECK();
    }
    return GPB_OK;
}

}  // namespace

extern "C" int gpb_system_size(const gpb_stack* st) {
    if (!st) return 0;
    return 3 * st->n_ori + st->n_rest + st->n_drift + st->n_faults;
}

// Systems up to this order use the single generic kernel (one launch; every reference example model), larger ones the
// blocked kernels above.
constexpr int kBlockedMinN = 512;

extern "C" int gpb_assemble_cov_ex(const gpb_stack* st, double* A, int lda, double* b, int flags, void* stream) {
    GPB_REQUIRE(st && A, "null argument");
    const int n = gpb_system_size(st);
    GPB_REQUIRE(n > 0 && lda >= n, "bad system size / lda");
    GPB_REQUIRE(st->n_drift == 0 || st->n_drift == 3 || st->n_drift == 9, "n_drift must be 0, 3 or 9");
    GPB_REQUIRE(st->n_faults == 0 || (st->fault_rest && st->fault_ref), "fault tables missing");
    GPB_REQUIRE(st->range > 0, "range must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    static const bool force_generic = getenv("GPB_COV_GENERIC") != nullptr;
    if (n >= kBlockedMinN && !force_generic) {
        const int lower_only = (flags & GPB_COV_LOWER_ONLY) ? 1 : 0;
        switch (st->kernel) {
            case GPB_KERNEL_CUBIC: return launch_blocked<GPB_KERNEL_CUBIC>(st, n, A, lda, b, lower_only, s);
            case GPB_KERNEL_EXPONENTIAL: return launch_blocked<GPB_KERNEL_EXPONENTIAL>(st, n, A, lda, b, lower_only, s);
            case GPB_KERNEL_MATERN52: return launch_blocked<GPB_KERNEL_MATERN52>(st, n, A, lda, b, lower_only, s);
            default: return gpb_set_error(GPB_E_INVALID, "unknown kernel function %d", st->kernel);
        }
    }
    const dim3 block(32, 8);
    const dim3 grid((n + kTile - 1) / kTile, (n + kTile - 1) / kTile);
    switch (st->kernel) {
        case GPB_KERNEL_CUBIC: cov_kernel<GPB_KERNEL_CUBIC><<<grid, block, 0, s>>>(*st, n, A, lda, b); break;
        case GPB_KERNEL_EXPONENTIAL: cov_kernel<GPB_KERNEL_EXPONENTIAL><<<grid, block, 0, s>>>(*st, n, A, lda, b); break;
        case GPB_KERNEL_MATERN52: cov_kernel<GPB_KERNEL_MATERN52><<<grid, block, 0, s>>>(*st, n, A, lda, b); break;
        default: return gpb_set_error(GPB_E_INVALID, "unknown kernel function %d", st->kernel);
    }
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_assemble_cov(const gpb_stack* st, double* A, int lda, double* b, void* stream) {
    return gpb_assemble_cov_ex(st, A, lda, b, 0, stream);
}
