// Marching cubes on the dense regular grid: the device-side replacement for the
// skimage.measure.marching_cubes call of gempy/modules/mesh_extranction/marching_cubes.py:82-89.
//
// HBM-bound byte/integer work.  Three passes over the lattice, 4096 consecutive points (z fastest) per CTA so that
// every load is coalesced and the +y/+x neighbours come out of L2:
//   (1) classify: which of the point's three +x/+y/+z edges carry a vertex, how many triangles its cube emits
//       -> one flag byte per point + per-CTA totals;
//   (2) a single-CTA scan of the per-CTA totals (vertices and triangles);
//   (3) emit vertices (and the per-point first-vertex id), then triangles (vertex ids through the owner point).
// Output order is deterministic: vertices by owner point then axis, triangles by cube then table order.
#include "gpb_common.cuh"
#include <mutex>

namespace {

constexpr int kBlock = 4096;         // lattice points per CTA
constexpr int kMaxTri = 5;

__constant__ unsigned char c_ntri[256];
__constant__ signed char c_tri[256 * kMaxTri * 3];

// edge id -> (lower corner, axis); corner id = 4*x + 2*y + z
__constant__ unsigned char c_edge_corner[12] = {0, 1, 2, 3, 0, 1, 4, 5, 0, 2, 4, 6};

// ---- case table, generated on the host ------------------------------------------------------------------
// On each face the crossings are joined so that every run of above-level corners is cut off by its own segment;
// segments are directed from the edge where a counter-clockwise walk (seen from outside) leaves the run to the edge
// where it entered; the closed loops are fanned from their lowest edge id.
struct CaseTable {
    unsigned char ntri[256];
    signed char tri[256 * kMaxTri * 3];
};

int edge_between(int a, int b) {
    static const int ec[12][2] = {{0, 4}, {1, 5}, {2, 6}, {3, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {0, 1}, {2, 3}, {4, 5}, {6, 7}};
    for (int e = 0; e < 12; ++e)
        if ((ec[e][0] == a && ec[e][1] == b) || (ec[e][0] == b && ec[e][1] == a)) return e;
    return -1;
}

bool build_case_table(CaseTable& t) {
    int faces[6][4];
    int nf = 0;
    for (int axis = 0; axis < 3; ++axis)
        for (int side = 0; side < 2; ++side) {
            int u = (axis + 1) % 3, v = (axis + 2) % 3;          // u x v = +axis
            if (side == 0) { const int s = u; u = v; v = s; }     // outward normal is -axis
            const int ab[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
            for (int q = 0; q < 4; ++q) {
                int c[3];
                c[axis] = side; c[u] = ab[q][0]; c[v] = ab[q][1];
                faces[nf][q] = 4 * c[0] + 2 * c[1] + c[2];
            }
            ++nf;
        }
    for (int cs = 0; cs < 256; ++cs) {
        int next[12];
        for (int e = 0; e < 12; ++e) next[e] = -1;
        auto in = [cs](int c) { return (cs >> c) & 1; };
        for (int f = 0; f < 6; ++f)
            for (int q = 0; q < 4; ++q) {
                const int prev = faces[f][(q + 3) & 3];
                if (!in(faces[f][q]) || in(prev)) continue;           // a run starts here
                const int entry = edge_between(prev, faces[f][q]);
                int r = q;
                while (in(faces[f][(r + 1) & 3])) ++r;
                const int leave = edge_between(faces[f][r & 3], faces[f][(r + 1) & 3]);
                next[leave] = entry;
            }
        bool seen[12] = {false};
        int n = 0;
        for (int e0 = 0; e0 < 12; ++e0) {
            if (next[e0] < 0 || seen[e0]) continue;
            int loop[12], len = 0;
            for (int e = e0; !seen[e]; e = next[e]) {
                seen[e] = true;
                loop[len++] = e;
                if (next[e] < 0) return false;
            }
            for (int k = 1; k + 1 < len; ++k) {
                if (n >= kMaxTri) return false;
                signed char* o = t.tri + (cs * kMaxTri + n) * 3;
                o[0] = (signed char)loop[0]; o[1] = (signed char)loop[k + 1]; o[2] = (signed char)loop[k];
                ++n;
            }
        }
        t.ntri[cs] = (unsigned char)n;
        for (int k = n; k < kMaxTri; ++k)
            for (int j = 0; j < 3; ++j) t.tri[(cs * kMaxTri + k) * 3 + j] = -1;
    }
    return true;
}

int upload_case_table() {
    static CaseTable table;
    static bool ok = false;
    static std::once_flag once;
    static std::mutex mu;
    static bool uploaded[64] = {false};
    std::call_once(once, [] { ok = build_case_table(table); });
    if (!ok) return gpb_set_error(GPB_E_INVALID, "marching-cubes case table generation failed");
    int dev = 0;
    GPB_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 64 && uploaded[dev]) return GPB_OK;
    GPB_CHECK_CUDA(cudaMemcpyToSymbol(c_ntri, table.ntri, sizeof(table.ntri)));
    GPB_CHECK_CUDA(cudaMemcpyToSymbol(c_tri, table.tri, sizeof(table.tri)));
    if (dev < 64) uploaded[dev] = true;
    return GPB_OK;
}

// ---- device helpers -----------------------------------------------------------------------------------------
// One CTA = kBlock consecutive lattice points.  The emit kernels run kThreads threads x kPts consecutive points each:
// flag bytes move as one 128-bit word per thread, first-vertex ids as 128-bit stores, and the CTA-wide scan runs over
// the per-thread partial sums.  The classify kernel (order-free: it only needs CTA totals) strides its kClassifyThreads
// threads over the CTA's points so that every access of a warp is contiguous.
constexpr int kPts = 16;
constexpr int kThreads = kBlock / kPts;
constexpr int kClassifyThreads = 256;

struct Lattice {
    int nx, ny, nz;
    long long m;
};

__device__ __forceinline__ void decompose(const Lattice& L, long long p, int& i, int& j, int& k) {
    if (L.m <= 0x7fffffffLL) {                       // 32-bit divisions (the 64-bit ones cost ~10x more)
        const unsigned pp = (unsigned)p;
        const unsigned r = pp / (unsigned)L.nz;
        k = (int)(pp - r * (unsigned)L.nz);
        const unsigned q = r / (unsigned)L.ny;
        j = (int)(r - q * (unsigned)L.ny);
        i = (int)q;
    } else {
        k = (int)(p % L.nz);
        const long long r = p / L.nz;
        j = (int)(r % L.ny);
        i = (int)(r / L.ny);
    }
}

// is the cube whose origin corner is (i, j, k) processed?  (inside the lattice and mask set at its far corner)
__device__ __forceinline__ bool cube_on(const Lattice& L, const unsigned char* __restrict__ mask, int i, int j, int k) {
    if (i < 0 || j < 0 || k < 0 || i >= L.nx - 1 || j >= L.ny - 1 || k >= L.nz - 1) return false;
    return mask == nullptr || mask[((long long)(i + 1) * L.ny + (j + 1)) * L.nz + (k + 1)] != 0;
}

// flag byte of lattice point p = (i, j, k): bits 0..2 = vertex on the +x/+y/+z edge, bits 3..5 = triangles of its cube
__device__ __forceinline__ unsigned classify(const Lattice& L, const double* __restrict__ Z, const unsigned char* __restrict__ mask,
                                             double level, long long p, int i, int j, int k) {
    const long long sy = L.nz, sx = (long long)L.ny * L.nz;
    const bool hx = i + 1 < L.nx, hy = j + 1 < L.ny, hz = k + 1 < L.nz;
    const bool s0 = Z[p] > level;
    unsigned flags = 0;
    bool c[8];
    c[0] = s0;
    c[4] = hx ? Z[p + sx] > level : s0;
    c[2] = hy ? Z[p + sy] > level : s0;
    c[1] = hz ? Z[p + 1] > level : s0;
    if (hx && c[4] != s0 &&
        (cube_on(L, mask, i, j - 1, k - 1) || cube_on(L, mask, i, j - 1, k) || cube_on(L, mask, i, j, k - 1) || cube_on(L, mask, i, j, k)))
        flags |= 1u;
    if (hy && c[2] != s0 &&
        (cube_on(L, mask, i - 1, j, k - 1) || cube_on(L, mask, i - 1, j, k) || cube_on(L, mask, i, j, k - 1) || cube_on(L, mask, i, j, k)))
        flags |= 2u;
    if (hz && c[1] != s0 &&
        (cube_on(L, mask, i - 1, j - 1, k) || cube_on(L, mask, i - 1, j, k) || cube_on(L, mask, i, j - 1, k) || cube_on(L, mask, i, j, k)))
        flags |= 4u;
    if (hx && hy && hz) {
        c[6] = Z[p + sx + sy] > level;
        c[5] = Z[p + sx + 1] > level;
        c[3] = Z[p + sy + 1] > level;
        c[7] = Z[p + sx + sy + 1] > level;
        unsigned cs = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) cs |= (c[q] ? 1u : 0u) << q;
        if (cs != 0u && cs != 255u && cube_on(L, mask, i, j, k)) flags |= (unsigned)c_ntri[cs] << 3;
    }
    return flags;
}

__device__ __forceinline__ int nverts_of(unsigned f) { return __popc(f & 7u); }
__device__ __forceinline__ int ntris_of(unsigned f) { return (int)(f >> 3); }

// the kPts flag bytes of a thread, as one 128-bit word when the run is complete and aligned
__device__ __forceinline__ void load_flags(const unsigned char* __restrict__ flags, long long p, long long m, unsigned f[kPts]) {
    static_assert(kPts == 16, "one uint4 of flag bytes per thread");
    if (p + kPts <= m && (reinterpret_cast<uintptr_t>(flags) & 15u) == 0) {
        const uint4 w4 = *reinterpret_cast<const uint4*>(flags + p);
        const unsigned w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int q = 0; q < kPts; ++q) f[q] = (w[q >> 2] >> (8 * (q & 3))) & 255u;
    } else {
#pragma unroll
        for (int q = 0; q < kPts; ++q) f[q] = (p + q < m) ? flags[p + q] : 0u;
    }
}

// CTA-wide exclusive scan of one int per thread (kThreads threads); returns the exclusive prefix
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums) {
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = (lane < kWarps) ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < kWarps; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kWarps) warp_sums[lane] = wi - w;                // exclusive over warps
    }
    __syncthreads();
    return warp_sums[warp] + incl - v;
}

__global__ void __launch_bounds__(kClassifyThreads) mc_classify_kernel(Lattice L, const double* __restrict__ Z, const unsigned char* __restrict__ mask,
                                                               double level, unsigned char* __restrict__ flags,
                                                               long long* __restrict__ block_counts, long long nblocks) {
    __shared__ int red[kClassifyThreads / 32];
    // coalesced mapping (only the CTA totals are needed here, so the order inside the CTA is free):
    // thread t takes points base + t, base + t + 256, ... -> every Z / mask / flag access of a warp is contiguous;
    // per-thread maxima 16 x 3 vertices and 16 x 5 triangles keep the packed halves apart up to the CTA total
    // (4096 x 3 and 4096 x 5 < 65536)
    long long p = (long long)blockIdx.x * kBlock + threadIdx.x;
    int packed = 0;                                   // vertices in the low half, triangles in the high half
    if (p < L.m) {
        int i, j, k;
        decompose(L, p, i, j, k);
#pragma unroll 4
        for (int q = 0; q < kBlock / kClassifyThreads; ++q) {
            if (p >= L.m) break;
            const unsigned f = classify(L, Z, mask, level, p, i, j, k);
            flags[p] = (unsigned char)f;
            packed += nverts_of(f) | (ntris_of(f) << 16);
            p += kClassifyThreads;
            k += kClassifyThreads;
            if (k >= L.nz) {
                const int c = k / L.nz;
                k -= c * L.nz;
                j += c;
                if (j >= L.ny) { const int d = j / L.ny; j -= d * L.ny; i += d; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) packed += __shfl_down_sync(0xffffffffu, packed, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = packed;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < kClassifyThreads / 32; ++w) t += red[w];
        block_counts[blockIdx.x] = t & 0xffff;
        block_counts[nblocks + 1 + blockIdx.x] = t >> 16;
    }
}

// exclusive scan of two arrays of nblocks entries (each followed by a total slot), one CTA,
// 16 consecutive entries per thread and 16 384 per sweep
__global__ void __launch_bounds__(1024) mc_scan_kernel(long long* counts, long long nblocks) {
    constexpr int kPer = 16;
    __shared__ long long wsum[32];
    __shared__ long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int which = 0; which < 2; ++which) {
        long long* a = counts + which * (nblocks + 1);
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (long long base = 0; base < nblocks; base += 1024 * kPer) {
            const long long i0 = base + (long long)threadIdx.x * kPer;
            long long v[kPer], own = 0;
#pragma unroll
            for (int e = 0; e < kPer; ++e) {
                v[e] = (i0 + e < nblocks) ? a[i0 + e] : 0;
                own += v[e];
            }
            long long incl = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const long long w = wsum[lane];
                long long wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long t = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += t;
                }
                wsum[lane] = wi - w;
            }
            __syncthreads();
            long long run = carry + wsum[warp] + incl - own;
#pragma unroll
            for (int e = 0; e < kPer; ++e) {
                if (i0 + e < nblocks) a[i0 + e] = run;
                run += v[e];
            }
            __syncthreads();
            if (threadIdx.x == 1023) carry = run;
            __syncthreads();
        }
        if (threadIdx.x == 0) a[nblocks] = carry;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) mc_vertices_kernel(Lattice L, const double* __restrict__ Z, const unsigned char* __restrict__ flags,
                                                               const long long* __restrict__ block_offsets, double level,
                                                               double ox, double oy, double oz, double dx, double dy, double dz,
                                                               int* __restrict__ vbase, double* __restrict__ vertices) {
    __shared__ int ws[kThreads / 32];
    const long long p = (long long)blockIdx.x * kBlock + (long long)threadIdx.x * kPts;
    unsigned f[kPts];
    load_flags(flags, p, L.m, f);
    int local[kPts], sum = 0;
#pragma unroll
    for (int q = 0; q < kPts; ++q) { local[q] = sum; sum += nverts_of(f[q]); }
    const int pre = block_exclusive_scan(sum, ws);
    if (p >= L.m) return;
    const long long first = block_offsets[blockIdx.x] + pre;
    if (p + kPts <= L.m && (reinterpret_cast<uintptr_t>(vbase) & 15u) == 0) {
#pragma unroll
        for (int q = 0; q < kPts; q += 4)
            *reinterpret_cast<int4*>(vbase + p + q) =
                make_int4((int)first + local[q], (int)first + local[q + 1], (int)first + local[q + 2], (int)first + local[q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < kPts; ++q)
            if (p + q < L.m) vbase[p + q] = (int)first + local[q];
    }
    if (sum == 0) return;
    const long long sx = (long long)L.ny * L.nz, sy = L.nz;
#pragma unroll
    for (int q = 0; q < kPts; ++q) {
        if ((f[q] & 7u) == 0) continue;
        int i, j, k;
        decompose(L, p + q, i, j, k);
        const double z0 = Z[p + q];
        long long slot = first + local[q];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (!(f[q] & (1u << a))) continue;
            const double z1 = Z[p + q + (a == 0 ? sx : (a == 1 ? sy : 1))];
            const double t = (level - z0) / (z1 - z0);
            double* o = vertices + 3 * slot;
            o[0] = ((double)i + (a == 0 ? t : 0.0)) * dx + ox;
            o[1] = ((double)j + (a == 1 ? t : 0.0)) * dy + oy;
            o[2] = ((double)k + (a == 2 ? t : 0.0)) * dz + oz;
            ++slot;
        }
    }
}

__global__ void __launch_bounds__(kThreads) mc_triangles_kernel(Lattice L, const double* __restrict__ Z, const unsigned char* __restrict__ flags,
                                                                const long long* __restrict__ block_offsets, double level,
                                                                const int* __restrict__ vbase, int* __restrict__ triangles) {
    __shared__ int ws[kThreads / 32];
    const long long p = (long long)blockIdx.x * kBlock + (long long)threadIdx.x * kPts;
    unsigned f[kPts];
    load_flags(flags, p, L.m, f);
    int local[kPts], sum = 0;
#pragma unroll
    for (int q = 0; q < kPts; ++q) { local[q] = sum; sum += ntris_of(f[q]); }
    const int pre = block_exclusive_scan(sum, ws);
    if (sum == 0) return;
    const long long sy = L.nz, sx = (long long)L.ny * L.nz;
    const long long first = block_offsets[blockIdx.x] + pre;
#pragma unroll
    for (int q = 0; q < kPts; ++q) {
        const int nt = ntris_of(f[q]);
        if (nt == 0) continue;
        const long long pq = p + q;
        unsigned cs = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const long long pc = pq + (c >> 2) * sx + ((c >> 1) & 1) * sy + (c & 1);
            cs |= (Z[pc] > level ? 1u : 0u) << c;
        }
        int* o = triangles + 3 * (first + local[q]);
        for (int t = 0; t < nt; ++t)
#pragma unroll
            for (int v = 0; v < 3; ++v) {
                const int e = c_tri[(cs * kMaxTri + t) * 3 + v];
                const int c = c_edge_corner[e];
                const int axis = e >> 2;
                const long long owner = pq + (c >> 2) * sx + ((c >> 1) & 1) * sy + (c & 1);
                const unsigned fo = flags[owner];
                o[3 * t + v] = vbase[owner] + __popc(fo & ((1u << axis) - 1u));
            }
    }
}

}  // namespace

extern "C" long long gpb_mc_scratch_elems(long long m) {
    const long long nblocks = (m + kBlock - 1) / kBlock;
    return 2 * (nblocks + 1);
}

extern "C" int gpb_mc_count(const double* Z, const unsigned char* mask, int nx, int ny, int nz, double level,
                            unsigned char* flags, long long* block_offsets, long long* n_vertices_host,
                            long long* n_triangles_host, void* stream) {
    GPB_REQUIRE(Z && flags && block_offsets && n_vertices_host && n_triangles_host, "null argument");
    GPB_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, "marching cubes needs at least 2 lattice points per axis");
    const int st = upload_case_table();
    if (st != GPB_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    Lattice L{nx, ny, nz, (long long)nx * ny * nz};
    const long long nblocks = (L.m + kBlock - 1) / kBlock;
    mc_classify_kernel<<<(unsigned)nblocks, kClassifyThreads, 0, s>>>(L, Z, mask, level, flags, block_offsets, nblocks);
    GPB_LAUNCH_CHECK();
    mc_scan_kernel<<<1, 1024, 0, s>>>(block_offsets, nblocks);
    GPB_LAUNCH_CHECK();
    long long totals[2] = {0, 0};
    GPB_CHECK_CUDA(cudaMemcpyAsync(&totals[0], block_offsets + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, s));
    GPB_CHECK_CUDA(cudaMemcpyAsync(&totals[1], block_offsets + 2 * nblocks + 1, sizeof(long long), cudaMemcpyDeviceToHost, s));
    GPB_CHECK_CUDA(cudaStreamSynchronize(s));
    if (totals[0] > 2147483647LL || totals[1] > 2147483647LL / 3)
        return gpb_set_error(GPB_E_INVALID, "mesh too large for 32-bit vertex ids: %lld vertices, %lld triangles", totals[0], totals[1]);
    *n_vertices_host = totals[0];
    *n_triangles_host = totals[1];
    return GPB_OK;
}

extern "C" int gpb_mc_emit(const double* Z, const unsigned char* flags, const long long* block_offsets, int nx, int ny,
                           int nz, double level, double ox, double oy, double oz, double dx, double dy, double dz,
                           int* vbase, double* vertices, int* triangles, void* stream) {
    GPB_REQUIRE(Z && flags && block_offsets && vbase, "null argument");
    GPB_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, "marching cubes needs at least 2 lattice points per axis");
    cudaStream_t s = (cudaStream_t)stream;
    Lattice L{nx, ny, nz, (long long)nx * ny * nz};
    const long long nblocks = (L.m + kBlock - 1) / kBlock;
    GPB_REQUIRE(vertices != nullptr, "vertices buffer is null");
    mc_vertices_kernel<<<(unsigned)nblocks, kThreads, 0, s>>>(L, Z, flags, block_offsets, level, ox, oy, oz, dx, dy, dz, vbase, vertices);
    GPB_LAUNCH_CHECK();
    if (triangles != nullptr) {
        mc_triangles_kernel<<<(unsigned)nblocks, kThreads, 0, s>>>(L, Z, flags, block_offsets + nblocks + 1, level, vbase, triangles);
        GPB_LAUNCH_CHECK();
    }
    return GPB_OK;
}
