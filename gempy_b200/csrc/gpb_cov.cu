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

}  // namespace

extern "C" int gpb_system_size(const gpb_stack* st) {
    if (!st) return 0;
    return 3 * st->n_ori + st->n_rest + st->n_drift + st->n_faults;
}

extern "C" int gpb_assemble_cov(const gpb_stack* st, double* A, int lda, double* b, void* stream) {
    GPB_REQUIRE(st && A, "null argument");
    const int n = gpb_system_size(st);
    GPB_REQUIRE(n > 0 && lda >= n, "bad system size / lda");
    GPB_REQUIRE(st->n_drift == 0 || st->n_drift == 3 || st->n_drift == 9, "n_drift must be 0, 3 or 9");
    GPB_REQUIRE(st->n_faults == 0 || (st->fault_rest && st->fault_ref), "fault tables missing");
    const dim3 block(32, 8);
    const dim3 grid((n + kTile - 1) / kTile, (n + kTile - 1) / kTile);
    cudaStream_t s = (cudaStream_t)stream;
    switch (st->kernel) {
        case GPB_KERNEL_CUBIC: cov_kernel<GPB_KERNEL_CUBIC><<<grid, block, 0, s>>>(*st, n, A, lda, b); break;
        case GPB_KERNEL_EXPONENTIAL: cov_kernel<GPB_KERNEL_EXPONENTIAL><<<grid, block, 0, s>>>(*st, n, A, lda, b); break;
        case GPB_KERNEL_MATERN52: cov_kernel<GPB_KERNEL_MATERN52><<<grid, block, 0, s>>>(*st, n, A, lda, b); break;
        default: return gpb_set_error(GPB_E_INVALID, "unknown kernel function %d", st->kernel);
    }
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}
