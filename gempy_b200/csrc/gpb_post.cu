// (4) Everything downstream of the scalar field: sigmoid activator, stack masks / combination, octree
// refinement test + child emission, dual-contouring edge crossings and per-voxel QEF vertices.
//
// Engine stages replaced (mask/combine, 4b, 4c), all reached from
// gempy_engine.compute_model (/root/reference/gempy/API/compute_API.py:68-73).  All of these are elementwise /
// compaction passes: HBM-bound, coalesced SoA accesses, grids sized from the element count.
// Semantics: oracle/gempy_oracle.py (activate, interpolate_all_fields masks, mark_voxels_by_corners,
// voxel_corners / voxel_children, edge_intersections, dual_contour_vertices).
#include "gpb_common.cuh"

namespace {

constexpr int kT = 256;
inline unsigned blocks_for(long long n, int per = kT) {
    long long b = (n + per - 1) / per;
    const long long cap = (long long)gpb_sm_count() * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ---- activator ------------------------------------------------------------------------------------------
constexpr int kMaxSurf = 64;

// 1 / (1 + exp(-x)); saturates exactly in FP64 beyond |x| ~ 40 / 745 (the slope is 5e6 by default, so almost
// every point takes one of the two early exits)
__device__ __forceinline__ double sigmoid(double x) {
    if (x > 40.0) return 1.0;
    if (x < -745.0) return 0.0;
    return 1.0 / (1.0 + exp(-x));
}

__global__ void activate_kernel(const double* __restrict__ Z, long long m, const double* __restrict__ iso,
                                const double* __restrict__ ids, int n, double slope, double* __restrict__ block) {
    __shared__ double s_iso[kMaxSurf], s_dif[kMaxSurf];
    __shared__ double s_base;
    if (threadIdx.x < n) {
        s_iso[threadIdx.x] = iso[threadIdx.x];
        s_dif[threadIdx.x] = ids[threadIdx.x] - ids[threadIdx.x + 1];
    }
    if (threadIdx.x == 0) s_base = ids[n];
    __syncthreads();
    // four independent loads in flight per thread (the kernel is a pure HBM stream: 8 B in, 8 B out per point)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < m; i0 += 4 * stride) {
        double z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long i = i0 + k * stride;
            z[k] = (i < m) ? Z[i] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long i = i0 + k * stride;
            if (i < m) {
                double v = s_base;
                for (int j = 0; j < n; ++j) v = fma(s_dif[j], sigmoid(slope * (z[k] - s_iso[j])), v);
                block[i] = v;
            }
        }
    }
}

// ---- min / shift ----------------------------------------------------------------------------------------
__global__ void set_inf_kernel(double* out) { *out = __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (v < __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

__global__ void min_kernel(const double* __restrict__ v, long long m, double* out) {
    double best = __longlong_as_double(0x7ff0000000000000LL);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x)
        best = fmin(best, v[i]);
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_down_sync(0xffffffffu, best, o));
    __shared__ double red[kT / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kT / 32; ++w) best = fmin(best, red[w]);
        atomic_min_double(out, best);
    }
}

__global__ void shift_kernel(const double* __restrict__ v, long long m, const double* __restrict__ minus, double* __restrict__ out) {
    const double s = *minus;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x)
        out[i] = v[i] - s;
}

// ---- combine --------------------------------------------------------------------------------------------
constexpr int kMaxStacks = 64;
struct CombineParams {
    int rel[kMaxStacks];
    int n;
};

__global__ void combine_kernel(const double* __restrict__ Z, const double* __restrict__ block, long long ld, long long m,
                               CombineParams cp, const double* __restrict__ iso_min, const double* __restrict__ iso_max,
                               double* __restrict__ final_block, double* __restrict__ faults_block,
                               unsigned char* __restrict__ squeezed, unsigned char* __restrict__ mask_out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long mask = 0ull;
        for (int s = 0; s < cp.n; ++s) {
            bool mk;
            switch (cp.rel[s]) {
                case GPB_REL_ERODE: mk = Z[s * ld + i] > iso_min[s]; break;
                case GPB_REL_ONLAP: mk = (s + 1 < cp.n) ? (Z[(s + 1) * ld + i] > iso_max[s + 1]) : true; break;
                case GPB_REL_FAULT: mk = false; break;
                default: mk = true;
            }
            if (mk) mask |= 1ull << s;
        }
        for (int s = cp.n - 2; s >= 0; --s)       // chained onlaps
            if (cp.rel[s] == GPB_REL_ONLAP && cp.rel[s + 1] == GPB_REL_ONLAP && !((mask >> (s + 1)) & 1ull))
                mask &= ~(1ull << s);
        bool free_ = true;
        double fin = 0.0, fau = 0.0;
        for (int s = 0; s < cp.n; ++s) {
            const bool mk = (mask >> s) & 1ull;
            const bool sq = mk && free_;
            free_ = free_ && !mk;
            const double b = block[s * ld + i];
            if (cp.rel[s] == GPB_REL_FAULT) fau += b;
            else if (sq) fin += b;
            squeezed[s * ld + i] = sq ? 1 : 0;
            if (mask_out) mask_out[s * ld + i] = mk ? 1 : 0;
        }
        final_block[i] = fin;
        faults_block[i] = fau;
    }
}

// ---- octree ---------------------------------------------------------------------------------------------
// sign pattern of corner / child c: x:----++++ y:--++--++ z:-+-+-+-+
__device__ __forceinline__ void sign3(int c, double& sx, double& sy, double& sz) {
    sx = (c & 4) ? 1.0 : -1.0;
    sy = (c & 2) ? 1.0 : -1.0;
    sz = (c & 1) ? 1.0 : -1.0;
}

__global__ void corners_kernel(const double* __restrict__ cen, long long ld_c, long long nvox, double hx, double hy, double hz,
                               double* __restrict__ out, long long ld_k) {
    const long long total = nvox * 8;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long v = e >> 3;
        double sx, sy, sz;
        sign3((int)(e & 7), sx, sy, sz);
        out[e] = cen[v] + sx * hx;
        out[ld_k + e] = cen[ld_c + v] + sy * hy;
        out[2 * ld_k + e] = cen[2 * ld_c + v] + sz * hz;
    }
}

__global__ void mark_kernel(const double* __restrict__ lith, const double* __restrict__ fault, long long nvox, int force_all,
                            unsigned char* __restrict__ mark) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
        bool differ = force_all != 0;
        if (!differ) {
            const double l0 = rint(lith[8 * v]);
            const double f0 = fault ? rint(fault[8 * v]) : 0.0;
            for (int c = 1; c < 8; ++c) {
                if (rint(lith[8 * v + c]) != l0) differ = true;
                if (fault && rint(fault[8 * v + c]) != f0) differ = true;
            }
        }
        mark[v] = differ ? 1 : 0;
    }
}

__global__ void rint_kernel(const double* __restrict__ in, long long n, double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = rint(in[i]);
}

__global__ void any8_kernel(const unsigned char* __restrict__ in, long long nvox, unsigned char* __restrict__ out) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
        unsigned w = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) w |= in[8 * v + c];
        out[v] = w ? 1 : 0;
    }
}

// block-wise exclusive scan of the marks (3 phases)
constexpr int kScanBlock = 1024;
__global__ void count_kernel(const unsigned char* __restrict__ mark, long long n, long long* __restrict__ counts) {
    __shared__ int red[kScanBlock / 32];
    const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
    int v = (i < n) ? (int)mark[i] : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kScanBlock / 32; ++w) t += red[w];
        counts[blockIdx.x] = t;
    }
}
__global__ void scan_counts_kernel(long long* counts, long long nblocks, long long* total) {
    // single thread block, sequential over chunks of 1024 (nblocks <= a few 10^4)
    __shared__ long long carry;
    __shared__ long long buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nblocks; base += 1024) {
        const long long i = base + threadIdx.x;
        buf[threadIdx.x] = (i < nblocks) ? counts[i] : 0;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {       // Hillis-Steele inclusive scan
            long long t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        const long long incl = buf[threadIdx.x];
        const long long own = (i < nblocks) ? counts[i] : 0;
        if (i < nblocks) counts[i] = carry + incl - own;       // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void emit_kernel(const double* __restrict__ cen, long long ld_c, long long nvox, const unsigned char* __restrict__ mark,
                            const long long* __restrict__ offsets, double qx, double qy, double qz,
                            double* __restrict__ out, long long ld_o) {
    __shared__ int warp_off[kScanBlock / 32];
    const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
    const int m = (i < nvox) ? (int)mark[i] : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, m);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_off[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < kScanBlock / 32; ++w) { const int t = warp_off[w]; warp_off[w] = acc; acc += t; }
    }
    __syncthreads();
    if (m) {
        const long long slot = offsets[blockIdx.x] + warp_off[warp] + __popc(ballot & ((1u << lane) - 1u));
        const double cx = cen[i], cy = cen[ld_c + i], cz = cen[2 * ld_c + i];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double sx, sy, sz;
            sign3(c, sx, sy, sz);
            out[8 * slot + c] = cx + sx * qx;
            out[ld_o + 8 * slot + c] = cy + sy * qy;
            out[2 * ld_o + 8 * slot + c] = cz + sz * qz;
        }
    }
}

// ---- forward gravity ---------------------------------------------------------------------------------------
// out[c] = sum_k tz[k] * density[id(c, k) - 1]; one CTA per device centre, fixed-order (deterministic) reduction
__global__ void __launch_bounds__(256) gravity_kernel(const double* __restrict__ block, const double* __restrict__ dens,
                                                      int n_dens, const double* __restrict__ tz, long long n_k,
                                                      double* __restrict__ out) {
    __shared__ double red[256];
    const double* b = block + (long long)blockIdx.x * n_k;
    double acc = 0.0;
    for (long long k = threadIdx.x; k < n_k; k += 256) {
        int id = (int)rint(b[k]);
        id = id < 1 ? 1 : (id > n_dens ? n_dens : id);
        acc = fma(dens[id - 1], tz[k], acc);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// ---- dual contouring -------------------------------------------------------------------------------------
__constant__ int c_edge_a[12] = {0, 1, 2, 3, 0, 1, 4, 5, 0, 2, 4, 6};
__constant__ int c_edge_b[12] = {4, 5, 6, 7, 2, 3, 6, 7, 1, 3, 5, 7};

__global__ void dc_edges_kernel(const double* __restrict__ cor, long long ld_k, const double* __restrict__ Zc, long long nvox,
                                double iso, const unsigned char* __restrict__ vmask, unsigned char* __restrict__ valid,
                                double* __restrict__ xyz) {
    const long long total = nvox * 12;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long v = e / 12;
        const int ed = (int)(e - v * 12);
        const long long ia = 8 * v + c_edge_a[ed], ib = 8 * v + c_edge_b[ed];
        const double za = Zc[ia], zb = Zc[ib];
        const double w = (iso - zb) / (za - zb);
        const bool ok = (w > 0.0) && (w < 1.0) && (!vmask || vmask[v]);
        valid[e] = ok ? 1 : 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double pa = cor[a * ld_k + ia], pb = cor[a * ld_k + ib];
            xyz[a * total + e] = ok ? pb + (pa - pb) * w : 0.0;
        }
    }
}

__global__ void dc_vertices_kernel(const unsigned char* __restrict__ valid, const double* __restrict__ xyz,
                                   const double* __restrict__ grad, long long nvox, double bias, double* __restrict__ vert) {
    const long long total = nvox * 12;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
        double M[6] = {0, 0, 0, 0, 0, 0};      // AtA: xx xy xz yy yz zz
        double r[3] = {0, 0, 0};                // Atb
        double msum[3] = {0, 0, 0};
        int mcnt[3] = {0, 0, 0};
        bool any = false;
        for (int ed = 0; ed < 12; ++ed) {
            const long long e = 12 * v + ed;
            if (!valid[e]) continue;
            any = true;
            const double p[3] = {xyz[e], xyz[total + e], xyz[2 * total + e]};
            const double n[3] = {grad[e], grad[total + e], grad[2 * total + e]};
            const double d = n[0] * p[0] + n[1] * p[1] + n[2] * p[2];
            M[0] += n[0] * n[0]; M[1] += n[0] * n[1]; M[2] += n[0] * n[2];
            M[3] += n[1] * n[1]; M[4] += n[1] * n[2]; M[5] += n[2] * n[2];
            r[0] += n[0] * d; r[1] += n[1] * d; r[2] += n[2] * d;
            for (int a = 0; a < 3; ++a)
                if (fabs(p[a]) > 1e-8) { msum[a] += p[a]; ++mcnt[a]; }      // np.isclose(x, 0) coordinates are ignored
        }
        if (!any) {
            vert[v] = nanv; vert[nvox + v] = nanv; vert[2 * nvox + v] = nanv;
            continue;
        }
        const double b2 = bias * bias;
        for (int a = 0; a < 3; ++a) {
            const double mass = msum[a] / (double)mcnt[a];       // 0/0 -> NaN, like nanmean of an empty slice
            r[a] += b2 * mass;
        }
        M[0] += b2; M[3] += b2; M[5] += b2;
        // symmetric 3x3 solve by the adjugate
        const double c00 = M[3] * M[5] - M[4] * M[4];
        const double c01 = M[2] * M[4] - M[1] * M[5];
        const double c02 = M[1] * M[4] - M[2] * M[3];
        const double c11 = M[0] * M[5] - M[2] * M[2];
        const double c12 = M[1] * M[2] - M[0] * M[4];
        const double c22 = M[0] * M[3] - M[1] * M[1];
        const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
        const double id = 1.0 / det;
        vert[v] = (c00 * r[0] + c01 * r[1] + c02 * r[2]) * id;
        vert[nvox + v] = (c01 * r[0] + c11 * r[1] + c12 * r[2]) * id;
        vert[2 * nvox + v] = (c02 * r[0] + c12 * r[1] + c22 * r[2]) * id;
    }
}

}  // namespace

extern "C" int gpb_activate(const double* Z, long long m, const double* isovalues, const double* ids, int n_surf,
                            double slope, double* block, void* stream) {
    GPB_REQUIRE(m >= 0 && n_surf >= 0 && n_surf <= kMaxSurf, "bad sizes (n_surf <= 64)");
    if (m == 0) return GPB_OK;
    GPB_REQUIRE(Z && ids && block && (n_surf == 0 || isovalues), "null argument");
    activate_kernel<<<blocks_for(m), kT, 0, (cudaStream_t)stream>>>(Z, m, isovalues, ids, n_surf, slope, block);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_min(const double* v, long long m, double* out_min, void* stream) {
    GPB_REQUIRE(v && out_min && m > 0, "bad arguments");
    set_inf_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out_min);
    GPB_LAUNCH_CHECK();
    min_kernel<<<blocks_for(m, kT * 8), kT, 0, (cudaStream_t)stream>>>(v, m, out_min);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_shift(const double* v, long long m, const double* minus, double* out, void* stream) {
    GPB_REQUIRE(v && minus && out && m >= 0, "bad arguments");
    if (m == 0) return GPB_OK;
    shift_kernel<<<blocks_for(m), kT, 0, (cudaStream_t)stream>>>(v, m, minus, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_combine(const double* Z, const double* block, long long ld, long long m, int n_stacks,
                           const int* relations_host, const double* iso_min, const double* iso_max, double* final_block,
                           double* faults_block, unsigned char* squeezed_mask, unsigned char* mask, void* stream) {
    GPB_REQUIRE(n_stacks >= 1 && n_stacks <= kMaxStacks, "1 <= n_stacks <= 64");
    GPB_REQUIRE(Z && block && relations_host && iso_min && iso_max && final_block && faults_block && squeezed_mask, "null argument");
    GPB_REQUIRE(m >= 0 && ld >= m, "bad sizes");
    if (m == 0) return GPB_OK;
    CombineParams cp;
    cp.n = n_stacks;
    for (int s = 0; s < n_stacks; ++s) cp.rel[s] = relations_host[s];
    combine_kernel<<<blocks_for(m), kT, 0, (cudaStream_t)stream>>>(Z, block, ld, m, cp, iso_min, iso_max, final_block,
                                                                   faults_block, squeezed_mask, mask);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_voxel_corners(const double* centers, long long ld_c, long long nvox, double hx, double hy, double hz,
                                 double* corners, long long ld_k, void* stream) {
    GPB_REQUIRE(centers && corners && nvox >= 0 && ld_c >= nvox && ld_k >= 8 * nvox, "bad arguments");
    if (nvox == 0) return GPB_OK;
    corners_kernel<<<blocks_for(nvox * 8), kT, 0, (cudaStream_t)stream>>>(centers, ld_c, nvox, hx, hy, hz, corners, ld_k);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_mark_voxels(const double* lith_corners, const double* fault_corners, long long nvox, int force_all,
                               unsigned char* mark, void* stream) {
    GPB_REQUIRE(lith_corners && mark && nvox >= 0, "bad arguments");
    if (nvox == 0) return GPB_OK;
    mark_kernel<<<blocks_for(nvox), kT, 0, (cudaStream_t)stream>>>(lith_corners, fault_corners, nvox, force_all, mark);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_emit_children(const double* centers, long long ld_c, long long nvox, const unsigned char* mark, double hx,
                                 double hy, double hz, double* children, long long ld_ch, long long* n_children_host,
                                 void* stream) {
    GPB_REQUIRE(centers && mark && n_children_host && nvox >= 0 && ld_c >= nvox, "bad arguments");
    *n_children_host = 0;
    if (nvox == 0) return GPB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const long long nblocks = (nvox + kScanBlock - 1) / kScanBlock;
    long long* counts = nullptr;
    GPB_CHECK_CUDA(gpb_malloc_async((void**)&counts, sizeof(long long) * (nblocks + 1), s));
    count_kernel<<<(unsigned)nblocks, kScanBlock, 0, s>>>(mark, nvox, counts);
    GPB_LAUNCH_CHECK();
    scan_counts_kernel<<<1, 1024, 0, s>>>(counts, nblocks, counts + nblocks);
    GPB_LAUNCH_CHECK();
    long long total = 0;
    GPB_CHECK_CUDA(cudaMemcpyAsync(&total, counts + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, s));
    GPB_CHECK_CUDA(cudaStreamSynchronize(s));
    *n_children_host = 8 * total;
    if (children != nullptr && total > 0) {
        if (ld_ch < 8 * total) {
            cudaFreeAsync(counts, s);
            return gpb_set_error(GPB_E_INVALID, "children buffer too small: need %lld columns, have %lld", 8 * total, ld_ch);
        }
        emit_kernel<<<(unsigned)nblocks, kScanBlock, 0, s>>>(centers, ld_c, nvox, mark, counts, hx, hy, hz, children, ld_ch);
        GPB_LAUNCH_CHECK();
    }
    GPB_CHECK_CUDA(cudaFreeAsync(counts, s));
    return GPB_OK;
}

// ---- octree -> regular fill: dense array at a level's resolution from the level above + the level's own voxels ----------
__global__ void upsample2_kernel(const double* __restrict__ src, int nx, int ny, int nz, double* __restrict__ dst) {
    const long long total = 8LL * nx * ny * nz;
    const int fy = 2 * ny, fz = 2 * nz;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / ((long long)fy * fz);
        const long long rem = e - i * fy * fz;
        const int j = (int)(rem / fz), k = (int)(rem - (long long)j * fz);
        dst[e] = src[((i >> 1) * ny + (j >> 1)) * nz + (k >> 1)];
    }
}

__global__ void scatter_lattice_kernel(const double* __restrict__ cen, long long ld_c, long long nvox, gpb_regular_grid g,
                                       const double* __restrict__ vals, int round_ids, double* __restrict__ dst) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
        const long long i = llrint((cen[v] - g.x0) / g.dx), j = llrint((cen[ld_c + v] - g.y0) / g.dy),
                        k = llrint((cen[2 * ld_c + v] - g.z0) / g.dz);
        if (i < 0 || j < 0 || k < 0 || i >= g.nx || j >= g.ny || k >= g.nz) continue;
        const double x = vals[v];
        dst[(i * g.ny + j) * g.nz + k] = round_ids ? rint(x) : x;
    }
}

extern "C" int gpb_upsample2(const double* src, int nx, int ny, int nz, double* dst, void* stream) {
    GPB_REQUIRE(src && dst && nx > 0 && ny > 0 && nz > 0, "bad arguments");
    upsample2_kernel<<<blocks_for(8LL * nx * ny * nz), kT, 0, (cudaStream_t)stream>>>(src, nx, ny, nz, dst);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_scatter_lattice(const double* centers, long long ld_c, long long nvox, const gpb_regular_grid* lattice,
                                   const double* vals, int round_ids, double* dst, void* stream) {
    GPB_REQUIRE(centers && lattice && vals && dst && nvox >= 0 && ld_c >= nvox, "bad arguments");
    if (nvox == 0) return GPB_OK;
    scatter_lattice_kernel<<<blocks_for(nvox), kT, 0, (cudaStream_t)stream>>>(centers, ld_c, nvox, *lattice, vals, round_ids, dst);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" long long gpb_scan_elems(long long nvox) { return (nvox + kScanBlock - 1) / kScanBlock + 1; }

extern "C" int gpb_count_marked(const unsigned char* mark, long long nvox, long long* offsets, long long* n_marked_host,
                                void* stream) {
    GPB_REQUIRE(mark && offsets && n_marked_host && nvox >= 0, "bad arguments");
    *n_marked_host = 0;
    if (nvox == 0) return GPB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const long long nblocks = (nvox + kScanBlock - 1) / kScanBlock;
    count_kernel<<<(unsigned)nblocks, kScanBlock, 0, s>>>(mark, nvox, offsets);
    GPB_LAUNCH_CHECK();
    scan_counts_kernel<<<1, 1024, 0, s>>>(offsets, nblocks, offsets + nblocks);
    GPB_LAUNCH_CHECK();
    GPB_CHECK_CUDA(cudaMemcpyAsync(n_marked_host, offsets + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, s));
    GPB_CHECK_CUDA(cudaStreamSynchronize(s));
    return GPB_OK;
}

extern "C" int gpb_emit_marked(const double* centers, long long ld_c, long long nvox, const unsigned char* mark,
                               const long long* offsets, double hx, double hy, double hz, double* children, long long ld_ch,
                               void* stream) {
    GPB_REQUIRE(centers && mark && offsets && children && nvox >= 0 && ld_c >= nvox, "bad arguments");
    if (nvox == 0) return GPB_OK;
    const long long nblocks = (nvox + kScanBlock - 1) / kScanBlock;
    emit_kernel<<<(unsigned)nblocks, kScanBlock, 0, (cudaStream_t)stream>>>(centers, ld_c, nvox, mark, offsets, hx, hy, hz, children, ld_ch);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_copy_2d(double* dst, long long ld_dst, const double* src, long long ld_src, long long rows, long long cols,
                           void* stream) {
    GPB_REQUIRE(dst && src && rows >= 0 && cols >= 0 && ld_dst >= cols && ld_src >= cols, "bad arguments");
    if (rows == 0 || cols == 0) return GPB_OK;
    GPB_CHECK_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * ld_dst, src, sizeof(double) * ld_src, sizeof(double) * cols, rows,
                                     cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return GPB_OK;
}

extern "C" int gpb_rint(const double* in, long long n, double* out, void* stream) {
    GPB_REQUIRE(in && out && n >= 0, "bad arguments");
    if (n == 0) return GPB_OK;
    rint_kernel<<<blocks_for(n), kT, 0, (cudaStream_t)stream>>>(in, n, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_any8(const unsigned char* in, long long nvox, unsigned char* out, void* stream) {
    GPB_REQUIRE(in && out && nvox >= 0, "bad arguments");
    if (nvox == 0) return GPB_OK;
    any8_kernel<<<blocks_for(nvox), kT, 0, (cudaStream_t)stream>>>(in, nvox, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_gravity(const double* block, const double* densities, int n_dens, const double* tz, int n_centers,
                           long long n_kernel, double* out, void* stream) {
    GPB_REQUIRE(block && densities && tz && out && n_dens >= 1 && n_centers >= 0 && n_kernel >= 1, "bad arguments");
    if (n_centers == 0) return GPB_OK;
    gravity_kernel<<<n_centers, 256, 0, (cudaStream_t)stream>>>(block, densities, n_dens, tz, n_kernel, out);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_dc_edges(const double* corners, long long ld_k, const double* Z_corners, long long nvox, double iso,
                            const unsigned char* voxel_mask, unsigned char* valid, double* xyz_edge, void* stream) {
    GPB_REQUIRE(corners && Z_corners && valid && xyz_edge && nvox >= 0 && ld_k >= 8 * nvox, "bad arguments");
    if (nvox == 0) return GPB_OK;
    dc_edges_kernel<<<blocks_for(nvox * 12), kT, 0, (cudaStream_t)stream>>>(corners, ld_k, Z_corners, nvox, iso, voxel_mask,
                                                                            valid, xyz_edge);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_dc_vertices(const unsigned char* valid, const double* xyz_edge, const double* grad_edge, long long nvox,
                               double bias, double* vertices, void* stream) {
    GPB_REQUIRE(valid && xyz_edge && grad_edge && vertices && nvox >= 0, "bad arguments");
    if (nvox == 0) return GPB_OK;
    dc_vertices_kernel<<<blocks_for(nvox), kT, 0, (cudaStream_t)stream>>>(valid, xyz_edge, grad_edge, nvox, bias, vertices);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}
