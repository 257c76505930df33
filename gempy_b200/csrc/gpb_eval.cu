// (3) Fused on-the-fly kernel-times-weights evaluation of the scalar field and its gradient.
//
// Engine stage replaced: "evaluator" + the evaluation kernels of "kernel_constructor" (SURVEY.md 8a2 row 3),
// reached from gempy_engine.compute_model (/root/reference/gempy/API/compute_API.py:68-73).  The reference
// design materialises an (n_data x chunk) kernel matrix per chunk (evaluation_chunk_size = 500000 elements,
// test/test_modules/test_serialize_model.*.verify/*.approved.txt) and multiplies it by the weights; here the
// kernel matrix never exists: every CTA streams the packed data-point table through shared memory with 1-D
// TMA bulk copies (cp.async.bulk + mbarrier, double buffered) and every thread keeps P grid points and their
// four accumulators (Z, dZ/dx, dZ/dy, dZ/dz) in registers.
//
// Roofline: FP64 pipe (DFMA).  Per (point, surface-point source): 23 FP64-pipe instructions + 1 MUFU;
// per (point, orientation): 33 + 2 MUFU (cubic kernel).  HBM traffic is the output only (32 B / point).
//
// Packed table (built by gpb_pack_eval_table), all coordinates divided by the range a:
//   [ n_sps_pad x {X, Y, Z, W} ]   W = c_o*i_res*w_i for rest points, -c_o*i_res*sum(w) for each reference point
//   [ n_ori_pad x {X, Y, Z, w'x, w'y, w'z} ]   w' = -(c_o*gi_res/a) * w
//   [ tail: mu_z[9] (drift, field), mu_g[9] (drift, gradient), scal[2], source moments S0, M1[3], M2, w_fault[n_faults] ]
#include "gpb_common.cuh"
#include <cstdlib>

// GPB_EVAL_SQRT3: third-order sqrt / reciprocal corrections (1e-20) instead of the second-order ones (2e-14 / 6e-14):
// one more FP64 instruction per pair each.  Measured on the 512^3 benchmark: see DESIGN.md section 5.
#ifdef GPB_EVAL_SQRT3
#define EVAL_SQRT gpb_fast_sqrt
#define EVAL_RCP gpb_fast_rcp
#else
#define EVAL_SQRT gpb_fast_sqrt2
#define EVAL_RCP gpb_fast_rcp2
#endif
#ifdef GPB_EVAL_LIBM_EXP
#define EVAL_EXP exp
#else
#define EVAL_EXP gpb_fast_exp_neg
#endif

namespace {

constexpr int kTileSp = 256;                  // sources per tile: 256 * 32 B = 8 KB (cubic: 256 * 80 B = 20 KB)
constexpr int kTileOri = 128;                 // 128 * 48 B = 6 KB
// Doubles per surface-point source record: {X, Y, Z, W} and, for the cubic kernel, {W p0, W p1, W p2, W q0, W q1, W q2},
// the coefficients of W P(u) and W Q(u) multiplied by the weight when the table is packed -- the pair loop then needs
// t u, W P(u) and one FMA for the field (one FP64 instruction per pair less than multiplying by W t in the loop).
__host__ __device__ constexpr int sp_rec(int kernel) { return kernel == GPB_KERNEL_CUBIC ? 10 : 4; }
static_assert(kTileSp * 4 * 8 >= kTileOri * 48, "the stage buffer must hold an orientation tile");
constexpr int kTailDoubles = 32;              // mu_z[9] mu_g[9] scal[2] moments S0, M1[3], M2 (+pad)

struct EvalParams {
    const double* src;       // packed table
    long long n_sps_pad;
    long long n_ori_pad;
    long long n_sps;         // actual number of surface-point sources (rest points + one per surface)
    long long n_ori;         // actual number of orientations
    int n_drift;
    int n_faults;
    // points
    gpb_regular_grid grid;   // REGULAR
    const double* xyz;       // !REGULAR: [3][ld_xyz]
    long long ld_xyz;
    long long i0;            // first global index (REGULAR)
    long long m;             // number of points (upper bound when m_dev is set)
    const long long* m_dev;  // optional (device): the actual number of points, read by the kernel (compacted lists)
    const double* fault_vals;   // row f of the fault values: fault_vals + (fault_ids ? fault_ids[f] : f) * ld_fault
    long long ld_fault;
    const int* fault_ids;       // optional (device): rows of the active faults inside a [n_stacks][ld] block matrix
    const double* fault_min;    // optional (device): per-row minimum subtracted from the fault values
    // fused activator (optional): block[k] = ids[n] + sum_j (ids[j] - ids[j+1]) sigma(slope (Z - iso[j]))
    double* block;
    const double* act_iso;
    const double* act_ids;
    int act_n;
    double act_slope;
    double* block_min;          // optional: running minimum of `block` over the launch (fault stacks)
    double* Z;
    double* gx;
    double* gy;
    double* gz;
    double inv_a;
    double eps_u;            // DIST_EPS / a^2
    double eps_reg;          // REG_EPS / a^2
};

// ---- covariance terms in range-normalised units (t = r/a, u = t^2) ------------------------------------
//   cval : C(r) (cubic: C - 1, the constant cancels because the source weights sum to zero)
//   kp   : a^2 * C'(r)/r
//   dd   : a^2 * (C'(r)/r - C''(r))          (numerator of the regularised gradient-gradient term)
// Cubic kernel, surface-point sources:  C - 1 = -7 u + t u P(u),   a^2 C'/r = -14 + t Q(u).
// The parts -7 u and -14 are polynomial in the point coordinates once summed over the sources
// (sum_s W_s u_s = (|X|^2 + eps) S0 - 2 X.M1 + M2,  sum_s W_s (X - s) = X S0 - M1), so they are added once per point
// from the source moments S0, M1, M2 (packed in the table tail) instead of once per pair: the pair loop accumulates
// only  Wt * (u P)  and  Wt * Q * d  with Wt = W t  -- one FP64 instruction less per pair.
//   returns c = u P(u) / t-free part to be multiplied by W t, kp = Q(u)
template <int KERNEL>
__device__ __forceinline__ void cov_sp(double u, double t, double& c, double& kp) {
    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {     // P(u) = 8.75 - 3.5 u + 0.75 u^2, Q(u) = 26.25 - 17.5 u + 5.25 u^2 (see sp_pair)
        const double P = fma(u, fma(0.75, u, -3.5), 8.75);
        kp = fma(u, fma(5.25, u, -17.5), 26.25);
        c = u * P;
    } else if constexpr (KERNEL == GPB_KERNEL_EXPONENTIAL) {
        const double e = EVAL_EXP(-0.5 * u);
        c = e;
        kp = -e;
    } else {
        const double s = 2.23606797749978969641 * t;
        const double e = EVAL_EXP(-s);
        c = fma(s, fma(s, 1.0 / 3.0, 1.0), 1.0) * e;
        kp = (-5.0 / 3.0) * (1.0 + s) * e;
    }
}

template <int KERNEL>
__device__ __forceinline__ void cov_ori(double u, double t, double& kp, double& dd) {
    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {
        const double Q = fma(u, fma(5.25, u, -17.5), 26.25);
        kp = fma(t, Q, -14.0);
        // dd = -(105/4) t (1-u)^2, written as -(s - s u)^2 t with s = sqrt(105/4)
        const double om = fma(-5.12347538297979853, u, 5.12347538297979853);
        dd = -(om * om) * t;
    } else if constexpr (KERNEL == GPB_KERNEL_EXPONENTIAL) {
        const double e = EVAL_EXP(-0.5 * u);
        kp = -e;
        dd = -e * u;
    } else {
        const double s = 2.23606797749978969641 * t;
        const double e = EVAL_EXP(-s);
        kp = (-5.0 / 3.0) * (1.0 + s) * e;
        dd = (-5.0 / 3.0) * e * s * s;
    }
}

// One surface-point pair: acc += W C(u) and (GRAD) g = W a^2 C'(r)/r, from the record's pre-multiplied coefficients (cubic)
// or its weight (other kernels).
template <int KERNEL> struct SpCoef {};
template <> struct SpCoef<GPB_KERNEL_CUBIC> { double2 c0, c1, c2; };     // {Wp0, Wp1}, {Wp2, Wq0}, {Wq1, Wq2}
template <int KERNEL>
__device__ __forceinline__ SpCoef<KERNEL> sp_coef(const double* rec) {
    SpCoef<KERNEL> cf;
    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {
        cf.c0 = *reinterpret_cast<const double2*>(rec + 4);
        cf.c1 = *reinterpret_cast<const double2*>(rec + 6);
        cf.c2 = *reinterpret_cast<const double2*>(rec + 8);
    }
    return cf;
}
template <int KERNEL, bool GRAD>
__device__ __forceinline__ void sp_pair(const SpCoef<KERNEL>& cf, double W, double u, double t, double& acc, double& g) {
    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {
        acc = fma(t * u, fma(u, fma(cf.c1.x, u, cf.c0.y), cf.c0.x), acc);
        if constexpr (GRAD) g = t * fma(u, fma(cf.c2.y, u, cf.c2.x), cf.c1.y);
    } else {
        double cv, kp;
        cov_sp<KERNEL>(u, t, cv, kp);
        acc = fma(W, cv, acc);
        if constexpr (GRAD) g = W * kp;
    }
}

// ---- fused activator + fault terms (epilogue helpers shared by both kernels) ----------------------------------------
constexpr int kActMax = 64;
struct ActShared {
    double iso[kActMax];
    double dif[kActMax];
    double base;
};

// 1 / (1 + exp(-x)); saturates exactly in FP64 beyond |x| ~ 40 / 745 (same expression as activate_kernel, gpb_post.cu)
__device__ __forceinline__ double act_sigmoid(double x) {
    if (x > 40.0) return 1.0;
    if (x < -745.0) return 0.0;
    return 1.0 / (1.0 + exp(-x));
}

__device__ __forceinline__ void act_load(const EvalParams& prm, ActShared& as) {
    if (prm.block != nullptr) {
        const int t = threadIdx.x;
        if (t < prm.act_n) {
            as.iso[t] = prm.act_iso[t];
            as.dif[t] = prm.act_ids[t] - prm.act_ids[t + 1];
        }
        if (t == 0) as.base = prm.act_ids[prm.act_n];
    }
}

__device__ __forceinline__ double act_value(const EvalParams& prm, const ActShared& as, double z) {
    double v = as.base;
    for (int j = 0; j < prm.act_n; ++j) v = fma(as.dif[j], act_sigmoid(prm.act_slope * (z - as.iso[j])), v);
    return v;
}

__device__ __forceinline__ double fault_term(const EvalParams& prm, const double* tail_faults, long long idx, double z) {
    for (int f = 0; f < prm.n_faults; ++f) {
        const long long row = prm.fault_ids ? prm.fault_ids[f] : f;
        double fv = prm.fault_vals[row * prm.ld_fault + idx];
        if (prm.fault_min) fv -= prm.fault_min[row];
        z = fma(tail_faults[f], fv, z);
    }
    return z;
}

__device__ __forceinline__ void act_min_commit(const EvalParams& prm, double vmin) {
    if (prm.block == nullptr || prm.block_min == nullptr) return;
    for (int o = 16; o > 0; o >>= 1) vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    if ((threadIdx.x & 31) == 0) {
        unsigned long long* a = reinterpret_cast<unsigned long long*>(prm.block_min);
        unsigned long long old = *a;
        while (vmin < __longlong_as_double((long long)old)) {
            const unsigned long long assumed = old;
            old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(vmin));
            if (old == assumed) break;
        }
    }
}

template <int KERNEL, bool GRAD, bool REGULAR, int P, int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
eval_kernel(const EvalParams prm) {
    constexpr int R = sp_rec(KERNEL);
    __shared__ __align__(128) double stage[2][kTileSp * R];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ ActShared act;
    double vmin = __longlong_as_double(0x7ff0000000000000LL);

    const int tid = threadIdx.x;
    act_load(prm, act);
    if (tid == 0) {
        gpb_mbar_init(&full[0], 1);
        gpb_mbar_init(&full[1], 1);
        gpb_fence_mbar_init();
    }
    __syncthreads();

    const long long n_sp_tiles = prm.n_sps_pad / kTileSp;
    const long long n_ori_tiles = prm.n_ori_pad / kTileOri;
    const long long n_tiles = n_sp_tiles + n_ori_tiles;
    const double* src_ori = prm.src + R * prm.n_sps_pad;
    const double* tail = src_ori + 6 * prm.n_ori_pad;

    const long long chunk = (long long)kThreads * P;
    const long long m_pts = (prm.m_dev != nullptr) ? min(prm.m, *prm.m_dev) : prm.m;
    const long long n_chunks = (m_pts + chunk - 1) / chunk;
    unsigned long long gt = 0;   // global tile counter of this CTA (buffer = gt & 1, parity = (gt >> 1) & 1)

    auto issue = [&](long long j, unsigned long long g) {
        // thread 0 only
        const int b = (int)(g & 1);
        if (j < n_sp_tiles) {
            gpb_mbar_expect_tx(&full[b], kTileSp * 8 * R);
            gpb_bulk_g2s(stage[b], prm.src + R * (long long)kTileSp * j, kTileSp * 8 * R, &full[b]);
        } else {
            gpb_mbar_expect_tx(&full[b], kTileOri * 48);
            gpb_bulk_g2s(stage[b], src_ori + 6 * (long long)kTileOri * (j - n_sp_tiles), kTileOri * 48, &full[b]);
        }
    };

    // Small tables (<= 2 tiles: every reference example model) are loaded once and stay resident in the two stage
    // buffers; larger ones are streamed per chunk, double buffered.
    const bool resident = n_tiles <= 2;
    if (resident && n_tiles > 0) {
        if (tid == 0) {
            issue(0, 0);
            if (n_tiles > 1) issue(1, 1);
        }
        gpb_mbar_wait(&full[0], 0);
        if (n_tiles > 1) gpb_mbar_wait(&full[1], 0);
    }

    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        // ---- this thread's P points -------------------------------------------------------------------
        double X[P], Y[P], Zc[P];
        double accZ[P], hx[P], hy[P], hz[P];
        long long idx[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            idx[k] = c * chunk + (long long)k * kThreads + tid;
            const long long i = idx[k] < m_pts ? idx[k] : m_pts - 1;     // clamp: tail threads recompute the last point
            double x, y, z;
            if constexpr (REGULAR) {
                const long long gi = prm.i0 + i;
                const long long nyz = (long long)prm.grid.ny * prm.grid.nz;
                const long long ix = gi / nyz;
                const long long rem = gi - ix * nyz;
                const long long iy = rem / prm.grid.nz;
                const long long iz = rem - iy * prm.grid.nz;
                x = fma((double)ix, prm.grid.dx, prm.grid.x0);
                y = fma((double)iy, prm.grid.dy, prm.grid.y0);
                z = fma((double)iz, prm.grid.dz, prm.grid.z0);
            } else {
                x = prm.xyz[i];
                y = prm.xyz[prm.ld_xyz + i];
                z = prm.xyz[2 * prm.ld_xyz + i];
            }
            X[k] = x * prm.inv_a;
            Y[k] = y * prm.inv_a;
            Zc[k] = z * prm.inv_a;
            accZ[k] = 0.0; hx[k] = 0.0; hy[k] = 0.0; hz[k] = 0.0;
        }

        if (!resident && tid == 0 && n_tiles > 0) issue(0, gt);

        for (long long j = 0; j < n_tiles; ++j, ++gt) {
            int b = (int)j;
            if (!resident) {
                if (tid == 0 && j + 1 < n_tiles) issue(j + 1, gt + 1);
                b = (int)(gt & 1);
                gpb_mbar_wait(&full[b], (uint32_t)((gt >> 1) & 1));
            }
            const double* s = stage[b];

            if (j < n_sp_tiles) {
                // ---- surface-point sources: {X, Y, Z, W} -----------------------------------------------
                const int cnt_sp = (int)min((long long)kTileSp, (prm.n_sps - j * kTileSp + 1) & ~1LL);
#pragma unroll 2
                for (int q = 0; q < cnt_sp; ++q) {
                    const double2 a0 = *reinterpret_cast<const double2*>(s + R * q);
                    const double2 a1 = *reinterpret_cast<const double2*>(s + R * q + 2);
                    const SpCoef<KERNEL> cf = sp_coef<KERNEL>(s + R * q);
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const double dx = X[k] - a0.x, dy = Y[k] - a0.y, dz = Zc[k] - a1.x;
                        const double u = fma(dz, dz, fma(dy, dy, fma(dx, dx, prm.eps_u)));
                        const double t = EVAL_SQRT(u);
                        double g;
                        sp_pair<KERNEL, GRAD>(cf, a1.y, u, t, accZ[k], g);
                        if constexpr (GRAD) {
                            hx[k] = fma(g, dx, hx[k]);
                            hy[k] = fma(g, dy, hy[k]);
                            hz[k] = fma(g, dz, hz[k]);
                        }
                    }
                }
                if (j == n_sp_tiles - 1) {
                    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {      // per-point part of the cubic split (see cov_sp)
                        const double S0 = tail[20], M1x = tail[21], M1y = tail[22], M1z = tail[23], M2 = tail[24];
#pragma unroll
                        for (int k = 0; k < P; ++k) {
                            const double x2 = fma(X[k], X[k], fma(Y[k], Y[k], fma(Zc[k], Zc[k], prm.eps_u)));
                            const double xm = fma(X[k], M1x, fma(Y[k], M1y, Zc[k] * M1z));
                            accZ[k] = fma(-7.0, fma(x2, S0, fma(-2.0, xm, M2)), accZ[k]);
                            if constexpr (GRAD) {
                                hx[k] = fma(-14.0, fma(X[k], S0, -M1x), hx[k]);
                                hy[k] = fma(-14.0, fma(Y[k], S0, -M1y), hy[k]);
                                hz[k] = fma(-14.0, fma(Zc[k], S0, -M1z), hz[k]);
                            }
                        }
                    }
                    if constexpr (GRAD) {
                        const double r = tail[18];       // gi^2 / i_res : surface-point part into H units
#pragma unroll
                        for (int k = 0; k < P; ++k) { hx[k] *= r; hy[k] *= r; hz[k] *= r; }
                    }
                }
            } else {
                // ---- orientation sources: {X, Y, Z, w'x, w'y, w'z} -------------------------------------
                const int cnt_or = (int)min((long long)kTileOri, (prm.n_ori - (j - n_sp_tiles) * kTileOri + 1) & ~1LL);
#pragma unroll 2
                for (int q = 0; q < cnt_or; ++q) {
                    const double2 a0 = *reinterpret_cast<const double2*>(s + 6 * q);
                    const double2 a1 = *reinterpret_cast<const double2*>(s + 6 * q + 2);
                    const double2 a2 = *reinterpret_cast<const double2*>(s + 6 * q + 4);
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const double dx = X[k] - a0.x, dy = Y[k] - a0.y, dz = Zc[k] - a1.x;
                        const double u = fma(dz, dz, fma(dy, dy, fma(dx, dx, prm.eps_u)));
                        const double t = EVAL_SQRT(u);
                        double kp, dd;
                        cov_ori<KERNEL>(u, t, kp, dd);
                        const double hw = fma(dz, a2.y, fma(dy, a2.x, dx * a1.y));
                        accZ[k] = fma(kp, hw, accZ[k]);
                        if constexpr (GRAD) {
                            const double c1 = -(dd * EVAL_RCP(u + prm.eps_reg)) * hw;
                            hx[k] = fma(c1, dx, fma(kp, a1.y, hx[k]));
                            hy[k] = fma(c1, dy, fma(kp, a2.x, hy[k]));
                            hz[k] = fma(c1, dz, fma(kp, a2.y, hz[k]));
                        }
                    }
                }
            }
            if (!resident) __syncthreads();     // everyone is done with stage[b] before thread 0 refills it
        }

        // ---- drift, faults, store -------------------------------------------------------------------------
        const double inv_agi = tail[19];        // 1 / (a * gi)
#pragma unroll
        for (int k = 0; k < P; ++k) {
            if (idx[k] >= m_pts) continue;
            double z = accZ[k];
            double g0 = hx[k] * inv_agi, g1 = hy[k] * inv_agi, g2 = hz[k] * inv_agi;
            if (prm.n_drift >= 3) {
                z = fma(tail[0], X[k], fma(tail[1], Y[k], fma(tail[2], Zc[k], z)));
                if constexpr (GRAD) { g0 += tail[9]; g1 += tail[10]; g2 += tail[11]; }
            }
            if (prm.n_drift == 9) {
                z = fma(tail[3], X[k] * X[k], fma(tail[4], Y[k] * Y[k], fma(tail[5], Zc[k] * Zc[k], z)));
                z = fma(tail[6], X[k] * Y[k], fma(tail[7], X[k] * Zc[k], fma(tail[8], Y[k] * Zc[k], z)));
                if constexpr (GRAD) {
                    g0 += 2.0 * tail[12] * X[k] + tail[15] * Y[k] + tail[16] * Zc[k];
                    g1 += 2.0 * tail[13] * Y[k] + tail[15] * X[k] + tail[17] * Zc[k];
                    g2 += 2.0 * tail[14] * Zc[k] + tail[16] * X[k] + tail[17] * Y[k];
                }
            }
            z = fault_term(prm, tail + kTailDoubles, idx[k], z);
            prm.Z[idx[k]] = z;
            if constexpr (GRAD) {
                prm.gx[idx[k]] = g0;
                prm.gy[idx[k]] = g1;
                prm.gz[idx[k]] = g2;
            }
            if (prm.block != nullptr) {
                const double v = act_value(prm, act, z);
                prm.block[idx[k]] = v;
                vmin = fmin(vmin, v);
            }
        }
    }
    act_min_commit(prm, vmin);
}

// ---- packing ----------------------------------------------------------------------------------------------
struct PackParams {
    gpb_stack st;
    const double* w;
    double* src;
    long long n_sps_pad, n_ori_pad;
};

__global__ void pack_kernel(const PackParams p) {
    const gpb_stack& st = p.st;
    const double inv_a = 1.0 / st.range;
    const double cI = st.c_o * st.i_res;
    const double cG = -(st.c_o * st.gi_res) * inv_a;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double* w_g = p.w;
    const double* w_i = p.w + 3LL * st.n_ori;
    const double* mu = w_i + st.n_rest;
    const double* w_f = mu + st.n_drift;
    const long long n_sps = (long long)st.n_rest + st.n_surf;
    const int rec = sp_rec(st.kernel);
    // surface-point sources
    for (long long i = tid; i < p.n_sps_pad; i += stride) {
        double x = 0, y = 0, z = 0, W = 0;
        if (i < st.n_rest) {
            x = st.rest[i]; y = st.rest[st.n_rest + i]; z = st.rest[2LL * st.n_rest + i];
            W = cI * w_i[i];
        } else if (i < n_sps) {
            const int s = (int)(i - st.n_rest);
            x = st.ref_unique[s]; y = st.ref_unique[st.n_surf + s]; z = st.ref_unique[2 * st.n_surf + s];
            double acc = 0.0;                       // sequential sum: deterministic
            for (int r = st.surf_offsets[s]; r < st.surf_offsets[s + 1]; ++r) acc += w_i[r];
            W = -cI * acc;
        }
        double* o = p.src + rec * i;
        o[0] = x * inv_a; o[1] = y * inv_a; o[2] = z * inv_a; o[3] = W;
        if (rec == 10) {
            o[4] = 8.75 * W; o[5] = -3.5 * W; o[6] = 0.75 * W;
            o[7] = 26.25 * W; o[8] = -17.5 * W; o[9] = 5.25 * W;
        }
    }
    // orientation sources
    double* so = p.src + rec * p.n_sps_pad;
    for (long long i = tid; i < p.n_ori_pad; i += stride) {
        double v[6] = {0, 0, 0, 0, 0, 0};
        if (i < st.n_ori) {
            v[0] = st.ori_pos[i] * inv_a; v[1] = st.ori_pos[st.n_ori + i] * inv_a; v[2] = st.ori_pos[2LL * st.n_ori + i] * inv_a;
            v[3] = cG * w_g[i]; v[4] = cG * w_g[st.n_ori + i]; v[5] = cG * w_g[2LL * st.n_ori + i];
        }
        for (int k = 0; k < 6; ++k) so[6 * i + k] = v[k];
    }
    // tail
    double* tl = so + 6 * p.n_ori_pad;
    if (tid == 0) {
        const double a = st.range;
        for (int k = 0; k < kTailDoubles; ++k) tl[k] = 0.0;
        for (int k = 0; k < st.n_drift; ++k) {
            const double m = mu[k];
            // field: gi * mu_k f_k(x), x = a X  ->  linear terms * a, quadratic * a^2
            tl[k] = st.gi_res * m * (k < 3 ? a : a * a);
            // gradient: mu_k d f_k / d x   (linear: mu; quadratic: coefficient of X is mu * a)
            tl[9 + k] = (k < 3) ? m : m * a;
        }
        tl[18] = st.gi_res * st.gi_res / st.i_res;
        tl[19] = 1.0 / (a * st.gi_res);
        // moments of the surface-point sources in the packed units (same W and X the pair loop sees)
        double S0 = 0.0, M1x = 0.0, M1y = 0.0, M1z = 0.0, M2 = 0.0;
        for (int i = 0; i < st.n_rest; ++i) {
            const double W = cI * w_i[i];
            const double x = st.rest[i] * inv_a, y = st.rest[st.n_rest + i] * inv_a, z = st.rest[2LL * st.n_rest + i] * inv_a;
            S0 += W; M1x = fma(W, x, M1x); M1y = fma(W, y, M1y); M1z = fma(W, z, M1z);
            M2 = fma(W, fma(x, x, fma(y, y, z * z)), M2);
        }
        for (int sidx = 0; sidx < st.n_surf; ++sidx) {
            double acc = 0.0;
            for (int r = st.surf_offsets[sidx]; r < st.surf_offsets[sidx + 1]; ++r) acc += w_i[r];
            const double W = -cI * acc;
            const double x = st.ref_unique[sidx] * inv_a, y = st.ref_unique[st.n_surf + sidx] * inv_a,
                         z = st.ref_unique[2 * st.n_surf + sidx] * inv_a;
            S0 += W; M1x = fma(W, x, M1x); M1y = fma(W, y, M1y); M1z = fma(W, z, M1z);
            M2 = fma(W, fma(x, x, fma(y, y, z * z)), M2);
        }
        tl[20] = S0; tl[21] = M1x; tl[22] = M1y; tl[23] = M1z; tl[24] = M2;
        for (int f = 0; f < st.n_faults; ++f) tl[kTailDoubles + f] = w_f[f];
    }
}


// ---- regular-grid "z-run" variant ------------------------------------------------------------------------
// Every thread owns P consecutive cells of one (x, y) column of the regular grid (z is the fastest axis,
// gempy/core/data/grid_modules/regular_grid.py:58-71).  dx, dy, dx^2 + dy^2 and (for orientations) dx w'x + dy w'y
// are computed once per (thread, source) and shared by the P points: 3 (surface points) / 4.5 (orientations)
// fewer FP64 instructions per pair at P = 4.  Requires nz % P == 0, i0 % P == 0, m % P == 0.
template <int KERNEL, bool GRAD, int P, int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
eval_zrun_kernel(const EvalParams prm) {
    constexpr int R = sp_rec(KERNEL);
    __shared__ __align__(128) double stage[2][kTileSp * R];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ ActShared act;
    double vmin = __longlong_as_double(0x7ff0000000000000LL);

    const int tid = threadIdx.x;
    act_load(prm, act);
    if (tid == 0) {
        gpb_mbar_init(&full[0], 1);
        gpb_mbar_init(&full[1], 1);
        gpb_fence_mbar_init();
    }
    __syncthreads();

    const long long n_sp_tiles = prm.n_sps_pad / kTileSp;
    const long long n_ori_tiles = prm.n_ori_pad / kTileOri;
    const long long n_tiles = n_sp_tiles + n_ori_tiles;
    const double* src_ori = prm.src + R * prm.n_sps_pad;
    const double* tail = src_ori + 6 * prm.n_ori_pad;

    const long long n_runs = prm.m / P;                       // runs of P points
    const long long n_chunks = (n_runs + kThreads - 1) / kThreads;
    unsigned long long gt = 0;

    auto issue = [&](long long j, unsigned long long g) {
        const int b = (int)(g & 1);
        if (j < n_sp_tiles) {
            gpb_mbar_expect_tx(&full[b], kTileSp * 8 * R);
            gpb_bulk_g2s(stage[b], prm.src + R * (long long)kTileSp * j, kTileSp * 8 * R, &full[b]);
        } else {
            gpb_mbar_expect_tx(&full[b], kTileOri * 48);
            gpb_bulk_g2s(stage[b], src_ori + 6 * (long long)kTileOri * (j - n_sp_tiles), kTileOri * 48, &full[b]);
        }
    };

    // Small tables (<= 2 tiles: every reference example model) are loaded once and stay resident in the two stage
    // buffers; larger ones are streamed per chunk, double buffered.
    const bool resident = n_tiles <= 2;
    if (resident && n_tiles > 0) {
        if (tid == 0) {
            issue(0, 0);
            if (n_tiles > 1) issue(1, 1);
        }
        gpb_mbar_wait(&full[0], 0);
        if (n_tiles > 1) gpb_mbar_wait(&full[1], 0);
    }

    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long run = c * kThreads + tid;
        const bool live = run < n_runs;
        const long long r = live ? run : n_runs - 1;            // idle tail threads recompute the last run
        const long long g0 = prm.i0 + r * P;
        const long long nyz = (long long)prm.grid.ny * prm.grid.nz;
        const long long ix = g0 / nyz;
        const long long rem = g0 - ix * nyz;
        const long long iy = rem / prm.grid.nz;
        const long long iz0 = rem - iy * prm.grid.nz;
        const double X = fma((double)ix, prm.grid.dx, prm.grid.x0) * prm.inv_a;
        const double Y = fma((double)iy, prm.grid.dy, prm.grid.y0) * prm.inv_a;
        double Zc[P], accZ[P], hx[P], hy[P], hz[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            Zc[k] = fma((double)(iz0 + k), prm.grid.dz, prm.grid.z0) * prm.inv_a;
            accZ[k] = 0.0; hx[k] = 0.0; hy[k] = 0.0; hz[k] = 0.0;
        }

        if (!resident && tid == 0 && n_tiles > 0) issue(0, gt);
        for (long long j = 0; j < n_tiles; ++j, ++gt) {
            int b = (int)j;
            if (!resident) {
                if (tid == 0 && j + 1 < n_tiles) issue(j + 1, gt + 1);
                b = (int)(gt & 1);
                gpb_mbar_wait(&full[b], (uint32_t)((gt >> 1) & 1));
            }
            const double* s = stage[b];
            if (j < n_sp_tiles) {
                const int cnt_sp = (int)min((long long)kTileSp, (prm.n_sps - j * kTileSp + 1) & ~1LL);
#pragma unroll 2
                for (int q = 0; q < cnt_sp; ++q) {
                    const double2 a0 = *reinterpret_cast<const double2*>(s + R * q);
                    const double2 a1 = *reinterpret_cast<const double2*>(s + R * q + 2);
                    const SpCoef<KERNEL> cf = sp_coef<KERNEL>(s + R * q);
                    const double dx = X - a0.x, dy = Y - a0.y;
                    const double pxy = fma(dy, dy, fma(dx, dx, prm.eps_u));
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const double dz = Zc[k] - a1.x;
                        const double u = fma(dz, dz, pxy);
                        const double t = EVAL_SQRT(u);
                        double g;
                        sp_pair<KERNEL, GRAD>(cf, a1.y, u, t, accZ[k], g);
                        if constexpr (GRAD) {
                            hx[k] = fma(g, dx, hx[k]);
                            hy[k] = fma(g, dy, hy[k]);
                            hz[k] = fma(g, dz, hz[k]);
                        }
                    }
                }
                if (j == n_sp_tiles - 1) {
                    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {      // per-point part of the cubic split (see cov_sp)
                        const double S0 = tail[20], M1x = tail[21], M1y = tail[22], M1z = tail[23], M2 = tail[24];
#pragma unroll
                        for (int k = 0; k < P; ++k) {
                            const double x2 = fma(X, X, fma(Y, Y, fma(Zc[k], Zc[k], prm.eps_u)));
                            const double xm = fma(X, M1x, fma(Y, M1y, Zc[k] * M1z));
                            accZ[k] = fma(-7.0, fma(x2, S0, fma(-2.0, xm, M2)), accZ[k]);
                            if constexpr (GRAD) {
                                hx[k] = fma(-14.0, fma(X, S0, -M1x), hx[k]);
                                hy[k] = fma(-14.0, fma(Y, S0, -M1y), hy[k]);
                                hz[k] = fma(-14.0, fma(Zc[k], S0, -M1z), hz[k]);
                            }
                        }
                    }
                    if constexpr (GRAD) {
                        const double rr = tail[18];
#pragma unroll
                        for (int k = 0; k < P; ++k) { hx[k] *= rr; hy[k] *= rr; hz[k] *= rr; }
                    }
                }
            } else {
                const int cnt_or = (int)min((long long)kTileOri, (prm.n_ori - (j - n_sp_tiles) * kTileOri + 1) & ~1LL);
#pragma unroll 2
                for (int q = 0; q < cnt_or; ++q) {
                    const double2 a0 = *reinterpret_cast<const double2*>(s + 6 * q);
                    const double2 a1 = *reinterpret_cast<const double2*>(s + 6 * q + 2);
                    const double2 a2 = *reinterpret_cast<const double2*>(s + 6 * q + 4);
                    const double dx = X - a0.x, dy = Y - a0.y;
                    const double pxy = fma(dy, dy, fma(dx, dx, prm.eps_u));
                    const double hwxy = fma(dy, a2.x, dx * a1.y);
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const double dz = Zc[k] - a1.x;
                        const double u = fma(dz, dz, pxy);
                        const double t = EVAL_SQRT(u);
                        double kp, dd;
                        cov_ori<KERNEL>(u, t, kp, dd);
                        const double hw = fma(dz, a2.y, hwxy);
                        accZ[k] = fma(kp, hw, accZ[k]);
                        if constexpr (GRAD) {
                            const double c1 = -(dd * EVAL_RCP(u + prm.eps_reg)) * hw;
                            hx[k] = fma(c1, dx, fma(kp, a1.y, hx[k]));
                            hy[k] = fma(c1, dy, fma(kp, a2.x, hy[k]));
                            hz[k] = fma(c1, dz, fma(kp, a2.y, hz[k]));
                        }
                    }
                }
            }
            if (!resident) __syncthreads();
        }

        const double inv_agi = tail[19];
        double zo[P], g0o[P], g1o[P], g2o[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            double z = accZ[k];
            double g0v = hx[k] * inv_agi, g1v = hy[k] * inv_agi, g2v = hz[k] * inv_agi;
            if (prm.n_drift >= 3) {
                z = fma(tail[0], X, fma(tail[1], Y, fma(tail[2], Zc[k], z)));
                if constexpr (GRAD) { g0v += tail[9]; g1v += tail[10]; g2v += tail[11]; }
            }
            if (prm.n_drift == 9) {
                z = fma(tail[3], X * X, fma(tail[4], Y * Y, fma(tail[5], Zc[k] * Zc[k], z)));
                z = fma(tail[6], X * Y, fma(tail[7], X * Zc[k], fma(tail[8], Y * Zc[k], z)));
                if constexpr (GRAD) {
                    g0v += 2.0 * tail[12] * X + tail[15] * Y + tail[16] * Zc[k];
                    g1v += 2.0 * tail[13] * Y + tail[15] * X + tail[17] * Zc[k];
                    g2v += 2.0 * tail[14] * Zc[k] + tail[16] * X + tail[17] * Y;
                }
            }
            if (live) z = fault_term(prm, tail + kTailDoubles, run * P + k, z);
            zo[k] = z; g0o[k] = g0v; g1o[k] = g1v; g2o[k] = g2v;
        }
        if (live) {
            const long long o = run * P;
            // P consecutive doubles per thread: 16-byte vector stores (alignment checked by the launcher)
#pragma unroll
            for (int k = 0; k < P; k += 2) {
                *reinterpret_cast<double2*>(prm.Z + o + k) = make_double2(zo[k], zo[k + 1]);
                if constexpr (GRAD) {
                    *reinterpret_cast<double2*>(prm.gx + o + k) = make_double2(g0o[k], g0o[k + 1]);
                    *reinterpret_cast<double2*>(prm.gy + o + k) = make_double2(g1o[k], g1o[k + 1]);
                    *reinterpret_cast<double2*>(prm.gz + o + k) = make_double2(g2o[k], g2o[k + 1]);
                }
            }
            if (prm.block != nullptr) {
#pragma unroll
                for (int k = 0; k < P; k += 2) {
                    const double v0 = act_value(prm, act, zo[k]), v1 = act_value(prm, act, zo[k + 1]);
                    *reinterpret_cast<double2*>(prm.block + o + k) = make_double2(v0, v1);
                    vmin = fmin(vmin, fmin(v0, v1));
                }
            }
        }
    }
    act_min_commit(prm, vmin);
}

// ---- octet variant: point lists made of complete sibling octets (the children of refined octree voxels) ---------------
// Points 8g .. 8g+7 are a parent centre +- a quarter cell in the pattern x:----++++ y:--++--++ z:-+-+-+-+
// (gpb_emit_marked / gpb_emit_children), so they take two x, two y and two z values.  Per (thread, source): six
// differences and four sums dx_a^2 + dy_b^2, then ONE FMA per point for the squared distance -- 2.5 instead of 6 FP64
// instructions per pair -- with the register tiling of the z-run kernel (8 points per thread).
template <int KERNEL, bool GRAD, int kThreads>
__global__ void __launch_bounds__(kThreads, 1)
eval_octet_kernel(const EvalParams prm) {
    constexpr int R = sp_rec(KERNEL);
    __shared__ __align__(128) double stage[2][kTileSp * R];
    __shared__ __align__(8) uint64_t full[2];
    __shared__ ActShared act;
    double vmin = __longlong_as_double(0x7ff0000000000000LL);

    const int tid = threadIdx.x;
    act_load(prm, act);
    if (tid == 0) {
        gpb_mbar_init(&full[0], 1);
        gpb_mbar_init(&full[1], 1);
        gpb_fence_mbar_init();
    }
    __syncthreads();

    const long long n_sp_tiles = prm.n_sps_pad / kTileSp;
    const long long n_ori_tiles = prm.n_ori_pad / kTileOri;
    const long long n_tiles = n_sp_tiles + n_ori_tiles;
    const double* src_ori = prm.src + R * prm.n_sps_pad;
    const double* tail = src_ori + 6 * prm.n_ori_pad;

    const long long n_oct = prm.m / 8;
    const long long n_chunks = (n_oct + kThreads - 1) / kThreads;
    unsigned long long gt = 0;

    auto issue = [&](long long j, unsigned long long g) {
        const int b = (int)(g & 1);
        if (j < n_sp_tiles) {
            gpb_mbar_expect_tx(&full[b], kTileSp * 8 * R);
            gpb_bulk_g2s(stage[b], prm.src + R * (long long)kTileSp * j, kTileSp * 8 * R, &full[b]);
        } else {
            gpb_mbar_expect_tx(&full[b], kTileOri * 48);
            gpb_bulk_g2s(stage[b], src_ori + 6 * (long long)kTileOri * (j - n_sp_tiles), kTileOri * 48, &full[b]);
        }
    };

    const bool resident = n_tiles <= 2;
    if (resident && n_tiles > 0) {
        if (tid == 0) {
            issue(0, 0);
            if (n_tiles > 1) issue(1, 1);
        }
        gpb_mbar_wait(&full[0], 0);
        if (n_tiles > 1) gpb_mbar_wait(&full[1], 0);
    }

    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long oct = c * kThreads + tid;
        const bool live = oct < n_oct;
        const long long base = 8 * (live ? oct : n_oct - 1);
        // the two values per axis, read from the points that carry them (point k: x index k >> 2, y (k >> 1) & 1, z k & 1)
        const double X[2] = {prm.xyz[base] * prm.inv_a, prm.xyz[base + 4] * prm.inv_a};
        const double Y[2] = {prm.xyz[prm.ld_xyz + base] * prm.inv_a, prm.xyz[prm.ld_xyz + base + 2] * prm.inv_a};
        const double Zc[2] = {prm.xyz[2 * prm.ld_xyz + base] * prm.inv_a, prm.xyz[2 * prm.ld_xyz + base + 1] * prm.inv_a};
        double accZ[8], hx[8], hy[8], hz[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { accZ[k] = 0.0; hx[k] = 0.0; hy[k] = 0.0; hz[k] = 0.0; }

        if (!resident && tid == 0 && n_tiles > 0) issue(0, gt);
        for (long long j = 0; j < n_tiles; ++j, ++gt) {
            int b = (int)j;
            if (!resident) {
                if (tid == 0 && j + 1 < n_tiles) issue(j + 1, gt + 1);
                b = (int)(gt & 1);
                gpb_mbar_wait(&full[b], (uint32_t)((gt >> 1) & 1));
            }
            const double* s = stage[b];
            if (j < n_sp_tiles) {
                const int cnt_sp = (int)min((long long)kTileSp, (prm.n_sps - j * kTileSp + 1) & ~1LL);
#pragma unroll 2
                for (int q = 0; q < cnt_sp; ++q) {
                    const double2 a0 = *reinterpret_cast<const double2*>(s + R * q);
                    const double2 a1 = *reinterpret_cast<const double2*>(s + R * q + 2);
                    const SpCoef<KERNEL> cf = sp_coef<KERNEL>(s + R * q);
                    const double dx[2] = {X[0] - a0.x, X[1] - a0.x}, dy[2] = {Y[0] - a0.y, Y[1] - a0.y};
                    const double dz[2] = {Zc[0] - a1.x, Zc[1] - a1.x};
                    const double px[2] = {fma(dx[0], dx[0], prm.eps_u), fma(dx[1], dx[1], prm.eps_u)};
                    const double pxy[4] = {fma(dy[0], dy[0], px[0]), fma(dy[1], dy[1], px[0]), fma(dy[0], dy[0], px[1]), fma(dy[1], dy[1], px[1])};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double u = fma(dz[k & 1], dz[k & 1], pxy[k >> 1]);
                        const double t = EVAL_SQRT(u);
                        double g;
                        sp_pair<KERNEL, GRAD>(cf, a1.y, u, t, accZ[k], g);
                        if constexpr (GRAD) {
                            hx[k] = fma(g, dx[k >> 2], hx[k]);
                            hy[k] = fma(g, dy[(k >> 1) & 1], hy[k]);
                            hz[k] = fma(g, dz[k & 1], hz[k]);
                        }
                    }
                }
                if (j == n_sp_tiles - 1) {
                    if constexpr (KERNEL == GPB_KERNEL_CUBIC) {
                        const double S0 = tail[20], M1x = tail[21], M1y = tail[22], M1z = tail[23], M2 = tail[24];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const double x = X[k >> 2], y = Y[(k >> 1) & 1], z = Zc[k & 1];
                            const double x2 = fma(x, x, fma(y, y, fma(z, z, prm.eps_u)));
                            const double xm = fma(x, M1x, fma(y, M1y, z * M1z));
                            accZ[k] = fma(-7.0, fma(x2, S0, fma(-2.0, xm, M2)), accZ[k]);
                            if constexpr (GRAD) {
                                hx[k] = fma(-14.0, fma(x, S0, -M1x), hx[k]);
                                hy[k] = fma(-14.0, fma(y, S0, -M1y), hy[k]);
                                hz[k] = fma(-14.0, fma(z, S0, -M1z), hz[k]);
                            }
                        }
                    }
                    if constexpr (GRAD) {
                        const double rr = tail[18];
#pragma unroll
                        for (int k = 0; k < 8; ++k) { hx[k] *= rr; hy[k] *= rr; hz[k] *= rr; }
                    }
                }
            } else {
                const int cnt_or = (int)min((long long)kTileOri, (prm.n_ori - (j - n_sp_tiles) * kTileOri + 1) & ~1LL);
#pragma unroll 2
                for (int q = 0; q < cnt_or; ++q) {
                    const double2 a0 = *reinterpret_cast<const double2*>(s + 6 * q);
                    const double2 a1 = *reinterpret_cast<const double2*>(s + 6 * q + 2);
                    const double2 a2 = *reinterpret_cast<const double2*>(s + 6 * q + 4);
                    const double dx[2] = {X[0] - a0.x, X[1] - a0.x}, dy[2] = {Y[0] - a0.y, Y[1] - a0.y};
                    const double dz[2] = {Zc[0] - a1.x, Zc[1] - a1.x};
                    const double px[2] = {fma(dx[0], dx[0], prm.eps_u), fma(dx[1], dx[1], prm.eps_u)};
                    const double pxy[4] = {fma(dy[0], dy[0], px[0]), fma(dy[1], dy[1], px[0]), fma(dy[0], dy[0], px[1]), fma(dy[1], dy[1], px[1])};
                    const double wx[2] = {dx[0] * a1.y, dx[1] * a1.y};
                    const double hwxy[4] = {fma(dy[0], a2.x, wx[0]), fma(dy[1], a2.x, wx[0]), fma(dy[0], a2.x, wx[1]), fma(dy[1], a2.x, wx[1])};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const double u = fma(dz[k & 1], dz[k & 1], pxy[k >> 1]);
                        const double t = EVAL_SQRT(u);
                        double kp, dd;
                        cov_ori<KERNEL>(u, t, kp, dd);
                        const double hw = fma(dz[k & 1], a2.y, hwxy[k >> 1]);
                        accZ[k] = fma(kp, hw, accZ[k]);
                        if constexpr (GRAD) {
                            const double c1 = -(dd * EVAL_RCP(u + prm.eps_reg)) * hw;
                            hx[k] = fma(c1, dx[k >> 2], fma(kp, a1.y, hx[k]));
                            hy[k] = fma(c1, dy[(k >> 1) & 1], fma(kp, a2.x, hy[k]));
                            hz[k] = fma(c1, dz[k & 1], fma(kp, a2.y, hz[k]));
                        }
                    }
                }
            }
            if (!resident) __syncthreads();
        }

        const double inv_agi = tail[19];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double x = X[k >> 2], y = Y[(k >> 1) & 1], zc = Zc[k & 1];
            double z = accZ[k];
            double g0v = hx[k] * inv_agi, g1v = hy[k] * inv_agi, g2v = hz[k] * inv_agi;
            if (prm.n_drift >= 3) {
                z = fma(tail[0], x, fma(tail[1], y, fma(tail[2], zc, z)));
                if constexpr (GRAD) { g0v += tail[9]; g1v += tail[10]; g2v += tail[11]; }
            }
            if (prm.n_drift == 9) {
                z = fma(tail[3], x * x, fma(tail[4], y * y, fma(tail[5], zc * zc, z)));
                z = fma(tail[6], x * y, fma(tail[7], x * zc, fma(tail[8], y * zc, z)));
                if constexpr (GRAD) {
                    g0v += 2.0 * tail[12] * x + tail[15] * y + tail[16] * zc;
                    g1v += 2.0 * tail[13] * y + tail[15] * x + tail[17] * zc;
                    g2v += 2.0 * tail[14] * zc + tail[16] * x + tail[17] * y;
                }
            }
            if (live) {
                const long long idx = base + k;
                z = fault_term(prm, tail + kTailDoubles, idx, z);
                prm.Z[idx] = z;
                if constexpr (GRAD) {
                    prm.gx[idx] = g0v;
                    prm.gy[idx] = g1v;
                    prm.gz[idx] = g2v;
                }
                if (prm.block != nullptr) {
                    const double v = act_value(prm, act, z);
                    prm.block[idx] = v;
                    vmin = fmin(vmin, v);
                }
            }
        }
    }
    act_min_commit(prm, vmin);
}

template <int KERNEL, bool GRAD>
int launch_octet(const EvalParams& prm, cudaStream_t stream) {
    constexpr int T = 256;
    const long long n_chunks = (prm.m / 8 + T - 1) / T;
    if (n_chunks == 0) return GPB_OK;
    long long grid = gpb_sm_count();
    if (grid > n_chunks) grid = n_chunks;
    eval_octet_kernel<KERNEL, GRAD, T><<<(unsigned)grid, T, 0, stream>>>(prm);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

template <int KERNEL, bool GRAD, int P, int T, int MINB>
int launch_zrun_cfg(const EvalParams& prm, cudaStream_t stream) {
    const long long n_runs = prm.m / P;
    const long long n_chunks = (n_runs + T - 1) / T;
    if (n_chunks == 0) return GPB_OK;
    int occ = 1;
    GPB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, eval_zrun_kernel<KERNEL, GRAD, P, T, MINB>, T, 0));
    if (occ < 1) occ = 1;
    long long grid = (long long)gpb_sm_count() * occ;
    if (grid > n_chunks) grid = n_chunks;
    eval_zrun_kernel<KERNEL, GRAD, P, T, MINB><<<(unsigned)grid, T, 0, stream>>>(prm);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

// z-run eligibility: whole runs of P cells inside one (x, y) column, 16-byte aligned outputs
template <int P>
bool zrun_ok(const EvalParams& prm) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return prm.grid.nz % P == 0 && prm.i0 % P == 0 && prm.m % P == 0 && prm.m >= P && al(prm.Z) &&
           (prm.gx == nullptr || (al(prm.gx) && al(prm.gy) && al(prm.gz))) && (prm.block == nullptr || al(prm.block));
}

template <int KERNEL, bool GRAD, bool REGULAR, int P, int T, int MINB>
int launch_eval_cfg(const EvalParams& prm, cudaStream_t stream) {
    const long long chunk = (long long)T * P;
    const long long n_chunks = (prm.m + chunk - 1) / chunk;
    if (n_chunks == 0) return GPB_OK;
    int occ = 1;
    GPB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, eval_kernel<KERNEL, GRAD, REGULAR, P, T, MINB>, T, 0));
    if (occ < 1) occ = 1;
    long long grid = (long long)gpb_sm_count() * occ;
    if (grid > n_chunks) grid = n_chunks;
    eval_kernel<KERNEL, GRAD, REGULAR, P, T, MINB><<<(unsigned)grid, T, 0, stream>>>(prm);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

// Launch configuration: P points per thread, T threads per CTA, MINB resident CTAs per SM.
// GPB_EVAL_VARIANT (environment, tuning only) selects alternative configurations of the cubic kernel.
template <int KERNEL, bool GRAD, bool REGULAR>
int launch_eval(const EvalParams& prm, cudaStream_t stream) {
    static const int variant = [] { const char* e = getenv("GPB_EVAL_VARIANT"); return e ? atoi(e) : 0; }();
    if constexpr (REGULAR) {
        // regular grids: z-run kernel whenever the range is made of whole runs (variant 100 forces the generic one)
        // measured on B200 (512^3-class grids, cubic, gradient): P=8/T=256 0.751 of the DFMA peak, P=8/T=128x2 0.734,
        // P=4/T=384 0.726, P=4/T=256 0.705, generic strided kernel 0.633
        // round 2, 16.5-instruction pair loop, 512^3: P=8/T=256 (226 registers) 779 ms; P=8 with T = 288 / 320 / 384 (ptxas
        // drops to 168 registers) 1108 / 993 / 826 ms; source loop unrolled by 1 / 4 instead of 2: 795 / 784 ms
        if (variant == 9 && zrun_ok<8>(prm)) return launch_zrun_cfg<KERNEL, GRAD, 8, 128, 2>(prm, stream);
        if (variant == 10 && zrun_ok<4>(prm)) return launch_zrun_cfg<KERNEL, GRAD, 4, 384, 1>(prm, stream);
        if (variant == 11 && zrun_ok<4>(prm)) return launch_zrun_cfg<KERNEL, GRAD, 4, 256, 1>(prm, stream);
        if (variant < 100 && zrun_ok<8>(prm)) return launch_zrun_cfg<KERNEL, GRAD, 8, 256, 1>(prm, stream);
        if (variant < 100 && zrun_ok<4>(prm)) return launch_zrun_cfg<KERNEL, GRAD, 4, 256, 1>(prm, stream);
    }
    // point lists (octree levels, corners, custom grids): P = 4 strided points per thread, 256 threads.  Measured on the
    // multi-fault octree-8 model (round 2): P = 8 / T = 256 41.6 ms, CTA sizes shrinking with the list length
    // (P = 2 / 1, T = 128 for the shallow levels) 42-45 ms, this configuration 35 ms -- the per-CTA set-up (table load,
    // barrier init) outweighs the better spread of tiny levels.
    return launch_eval_cfg<KERNEL, GRAD, REGULAR, 4, 256, 1>(prm, stream);
}

int dispatch_octet(int kernel, bool grad, const EvalParams& prm, cudaStream_t stream) {
#define GPB_OCT(K) case K: return grad ? launch_octet<K, true>(prm, stream) : launch_octet<K, false>(prm, stream);
    switch (kernel) {
        GPB_OCT(GPB_KERNEL_CUBIC)
        GPB_OCT(GPB_KERNEL_EXPONENTIAL)
        GPB_OCT(GPB_KERNEL_MATERN52)
        default: return gpb_set_error(GPB_E_INVALID, "unknown kernel function %d", kernel);
    }
#undef GPB_OCT
}

template <bool REGULAR>
int dispatch_eval(int kernel, bool grad, const EvalParams& prm, cudaStream_t stream) {
#define GPB_CASE(K)                                                         \
    case K:                                                                 \
        return grad ? launch_eval<K, true, REGULAR>(prm, stream) : launch_eval<K, false, REGULAR>(prm, stream);
    switch (kernel) {
        GPB_CASE(GPB_KERNEL_CUBIC)
        GPB_CASE(GPB_KERNEL_EXPONENTIAL)
        GPB_CASE(GPB_KERNEL_MATERN52)
        default: return gpb_set_error(GPB_E_INVALID, "unknown kernel function %d", kernel);
    }
#undef GPB_CASE
}

int fill_common(const gpb_stack* st, const double* src, EvalParams& prm) {
    GPB_REQUIRE(st != nullptr && src != nullptr, "null stack or table");
    GPB_REQUIRE(st->range > 0, "range must be positive");
    prm.src = src;
    prm.n_sps_pad = gpb_round_up((long long)st->n_rest + st->n_surf, kTileSp);
    prm.n_ori_pad = gpb_round_up(st->n_ori, kTileOri);
    prm.n_sps = (long long)st->n_rest + st->n_surf;
    prm.n_ori = st->n_ori;
    prm.n_drift = st->n_drift;
    prm.n_faults = st->n_faults;
    prm.inv_a = 1.0 / st->range;
    prm.eps_u = GPB_DIST_EPS / (st->range * st->range);
    prm.eps_reg = GPB_REG_EPS / (st->range * st->range);
    return GPB_OK;
}

}  // namespace

extern "C" long long gpb_eval_table_doubles(const gpb_stack* st) {
    if (!st) return 0;
    return sp_rec(st->kernel) * gpb_round_up((long long)st->n_rest + st->n_surf, kTileSp) + 6 * gpb_round_up(st->n_ori, kTileOri) +
           kTailDoubles + st->n_faults + 8;
}

extern "C" int gpb_pack_eval_table(const gpb_stack* st, const double* w, double* src, void* stream) {
    GPB_REQUIRE(st && w && src, "null argument");
    GPB_REQUIRE(st->n_drift == 0 || st->n_drift == 3 || st->n_drift == 9, "n_drift must be 0, 3 or 9");
    PackParams p;
    p.st = *st;
    p.w = w;
    p.src = src;
    p.n_sps_pad = gpb_round_up((long long)st->n_rest + st->n_surf, kTileSp);
    p.n_ori_pad = gpb_round_up(st->n_ori, kTileOri);
    const long long work = p.n_sps_pad > p.n_ori_pad ? p.n_sps_pad : p.n_ori_pad;
    int blocks = (int)((work + 127) / 128);
    if (blocks < 1) blocks = 1;
    pack_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(p);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

// Internal entry shared with the level executor (gpb_model.cu): one evaluation launch described by a GpbEvalCall.
int gpb_eval_call(const GpbEvalCall& c, cudaStream_t stream) {
    EvalParams prm{};
    int rc = fill_common(c.st, c.src, prm);
    if (rc) return rc;
    GPB_REQUIRE(c.m >= 0, "bad point count");
    if (c.m == 0) return GPB_OK;
    GPB_REQUIRE(c.Z != nullptr, "null output");
    GPB_REQUIRE((c.gx == nullptr) == (c.gy == nullptr) && (c.gx == nullptr) == (c.gz == nullptr), "gradient outputs: all or none");
    GPB_REQUIRE(c.st->n_faults == 0 || c.fault_vals != nullptr, "fault values missing");
    GPB_REQUIRE(c.block == nullptr || (c.act_ids != nullptr && c.act_n >= 0 && c.act_n <= kActMax && (c.act_n == 0 || c.act_iso != nullptr)),
                "fused activator: ids / isovalues missing or more than 64 surfaces");
    if (c.regular) {
        GPB_REQUIRE(c.i0 >= 0 && c.i0 + c.m <= (long long)c.grid.nx * c.grid.ny * c.grid.nz, "bad point range");
        prm.grid = c.grid;
        prm.i0 = c.i0;
    } else {
        GPB_REQUIRE(c.xyz != nullptr && c.ld_xyz >= c.m, "null points / bad leading dimension");
        prm.xyz = c.xyz;
        prm.ld_xyz = c.ld_xyz;
    }
    prm.m = c.m;
    prm.fault_vals = c.fault_vals;
    prm.ld_fault = c.ld_fault;
    prm.fault_ids = c.fault_ids;
    prm.fault_min = c.fault_min;
    prm.block = c.block;
    prm.act_iso = c.act_iso;
    prm.act_ids = c.act_ids;
    prm.act_n = c.act_n;
    prm.act_slope = c.act_slope;
    prm.block_min = c.block_min;
    prm.m_dev = c.regular ? nullptr : c.m_dev;
    prm.Z = c.Z; prm.gx = c.gx; prm.gy = c.gy; prm.gz = c.gz;
    static const bool no_octets = getenv("GPB_NO_OCTETS") != nullptr;
    // complete sibling octets, enough of them to fill the machine: the octet kernel (8 points per thread)
    if (!c.regular && c.octets && !no_octets && c.m_dev == nullptr && c.m % 8 == 0 && c.m >= 8LL * 256 * 64)
        return dispatch_octet(c.st->kernel, c.gx != nullptr, prm, stream);
    return c.regular ? dispatch_eval<true>(c.st->kernel, c.gx != nullptr, prm, stream)
                     : dispatch_eval<false>(c.st->kernel, c.gx != nullptr, prm, stream);
}

extern "C" int gpb_eval_regular(const gpb_stack* st, const double* src, const gpb_regular_grid* grid, long long i0,
                                long long i1, const double* fault_vals, long long ld_fault, double* Z, double* gx,
                                double* gy, double* gz, void* stream) {
    GPB_REQUIRE(grid && Z, "null grid or output");
    GPB_REQUIRE(i0 >= 0 && i1 >= i0 && i1 <= (long long)grid->nx * grid->ny * grid->nz, "bad point range");
    GpbEvalCall c;
    c.st = st; c.src = src;
    c.regular = 1; c.grid = *grid; c.i0 = i0; c.m = i1 - i0;
    c.fault_vals = fault_vals; c.ld_fault = ld_fault;
    c.Z = Z; c.gx = gx; c.gy = gy; c.gz = gz;
    return gpb_eval_call(c, (cudaStream_t)stream);
}

extern "C" int gpb_eval_points(const gpb_stack* st, const double* src, const double* xyz, long long ld_xyz, long long m,
                               const double* fault_vals, long long ld_fault, double* Z, double* gx, double* gy,
                               double* gz, void* stream) {
    GPB_REQUIRE(m >= 0 && ld_xyz >= m, "bad point count");
    if (m == 0) return GPB_OK;
    GPB_REQUIRE(xyz && Z, "null points or output");
    GpbEvalCall c;
    c.st = st; c.src = src;
    c.xyz = xyz; c.ld_xyz = ld_xyz; c.m = m;
    c.fault_vals = fault_vals; c.ld_fault = ld_fault;
    c.Z = Z; c.gx = gx; c.gy = gy; c.gz = gz;
    return gpb_eval_call(c, (cudaStream_t)stream);
}
