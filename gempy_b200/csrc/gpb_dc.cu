// (4c) Dual contouring of one surface, entirely on the device.
//
// Engine stage replaced: "dual_contouring" (call site /root/reference/gempy/API/compute_API.py:68-73;
// consumer gempy/core/data/geo_model.py:110-121 reads dc_meshes[e].vertices / .edges).  Round 1 evaluated the gradient at
// all 12 n_vox edge slots, triangulated on the host with numpy and copied four arrays per surface over PCIe.  Here:
//
//   1. edge crossings of the isovalue on the 12 edges of every voxel (valid flags + crossing points)
//   2. stable compaction of the valid crossings (block counts -> scan -> scatter), typically < 10 % of the slots
//   3. gradient of the stack's field at the COMPACTED crossings only (fused evaluation kernel; the number of points is read
//      from device memory, so no host synchronisation is needed to size the launch)
//   4. one QEF vertex per voxel that has a crossing (12 edge planes + 3 mass-point planes), compacted in voxel order
//   5. triangulation: a hash table voxel lattice code -> vertex id; every crossed edge that is shared by four surface
//      voxels (edges 3, 7, 11 = the voxel's +y+z / +x+z / +x+y corner edges) emits two triangles, in (axis, voxel) order
//
// Output order is deterministic and equal to the oracle's (vertices in voxel order, triangles axis-major), so meshes compare
// exactly.  Nothing is copied to the host: the caller sizes the buffers for the worst case and reads the three counts
// (crossings, vertices, triangles) when somebody asks for the mesh.
#include "gpb_common.cuh"

namespace {

constexpr int kT = 256;
constexpr int kScanB = 1024;

inline unsigned grid_for(long long n, int per = kT) {
    long long b = (n + per - 1) / per;
    const long long cap = (long long)gpb_sm_count() * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

__constant__ int d_edge_a[12] = {0, 1, 2, 3, 0, 1, 4, 5, 0, 2, 4, 6};
__constant__ int d_edge_b[12] = {4, 5, 6, 7, 2, 3, 6, 7, 1, 3, 5, 7};

// voxel ownership: any corner inside the stack's squeezed mask (NULL mask: every voxel)
__global__ void dc_edges2_kernel(const double* __restrict__ cor, long long ld_k, const double* __restrict__ Zc, long long nvox,
                                 const double* __restrict__ iso_dev, const unsigned char* __restrict__ sq_corners,
                                 unsigned char* __restrict__ valid, double* __restrict__ xyz) {
    const long long total = nvox * 12;
    const double iso = *iso_dev;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long v = e / 12;
        const int ed = (int)(e - v * 12);
        const long long ia = 8 * v + d_edge_a[ed], ib = 8 * v + d_edge_b[ed];
        const double za = Zc[ia], zb = Zc[ib];
        const double w = (iso - zb) / (za - zb);
        bool own = true;
        if (sq_corners) {
            unsigned any = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) any |= sq_corners[8 * v + c];
            own = any != 0;
        }
        const bool ok = (w > 0.0) && (w < 1.0) && own;
        valid[e] = ok ? 1 : 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double pa = cor[a * ld_k + ia], pb = cor[a * ld_k + ib];
            xyz[a * total + e] = ok ? pb + (pa - pb) * w : 0.0;
        }
    }
}

// ---- stable compaction helpers: counts per block of 1024 flags, exclusive scan of the counts, rank inside the block ----
__global__ void block_count_kernel(const unsigned char* __restrict__ flag, long long n, long long* __restrict__ counts) {
    __shared__ int red[kScanB / 32];
    const long long i = (long long)blockIdx.x * kScanB + threadIdx.x;
    int v = (i < n) ? (int)(flag[i] != 0) : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kScanB / 32; ++w) t += red[w];
        counts[blockIdx.x] = t;
    }
}

__global__ void scan_blocks_kernel(long long* counts, long long nblocks, long long* total) {
    __shared__ long long carry;
    __shared__ long long buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nblocks; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long own = (i < nblocks) ? counts[i] : 0;
        buf[threadIdx.x] = own;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const long long t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        const long long incl = buf[threadIdx.x];
        if (i < nblocks) counts[i] = carry + incl - own;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// rank of element i among the set flags (valid inside a block of kScanB threads; every thread of the block must call)
__device__ __forceinline__ long long block_rank(int m, const long long* __restrict__ offsets, int* warp_off) {
    const unsigned ballot = __ballot_sync(0xffffffffu, m);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_off[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < kScanB / 32; ++w) { const int t = warp_off[w]; warp_off[w] = acc; acc += t; }
    }
    __syncthreads();
    return offsets[blockIdx.x] + warp_off[warp] + __popc(ballot & ((1u << lane) - 1u));
}

__global__ void compact_edges_kernel(const unsigned char* __restrict__ valid, long long total, const long long* __restrict__ offsets,
                                     const double* __restrict__ xyz, int* __restrict__ pos, double* __restrict__ xyz_c, long long ld_c) {
    __shared__ int warp_off[kScanB / 32];
    const long long e = (long long)blockIdx.x * kScanB + threadIdx.x;
    const int m = (e < total) ? (int)valid[e] : 0;
    const long long p = block_rank(m, offsets, warp_off);
    if (e < total) pos[e] = m ? (int)p : -1;
    if (m) {
        xyz_c[p] = xyz[e];
        xyz_c[ld_c + p] = xyz[total + e];
        xyz_c[2 * ld_c + p] = xyz[2 * total + e];
    }
}

// Per-voxel QEF (12 edge planes with raw gradient normals + 3 mass-point planes of strength `bias`), inputs compacted
__global__ void dc_vertices2_kernel(const unsigned char* __restrict__ valid, const int* __restrict__ pos, const double* __restrict__ xyz_c,
                                    const double* __restrict__ grad_c, long long ld_c, long long nvox, double bias,
                                    double* __restrict__ vert, unsigned char* __restrict__ has_v) {
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
            const long long q = pos[e];
            const double p[3] = {xyz_c[q], xyz_c[ld_c + q], xyz_c[2 * ld_c + q]};
            const double n[3] = {grad_c[q], grad_c[ld_c + q], grad_c[2 * ld_c + q]};
            const double d = n[0] * p[0] + n[1] * p[1] + n[2] * p[2];
            M[0] += n[0] * n[0]; M[1] += n[0] * n[1]; M[2] += n[0] * n[2];
            M[3] += n[1] * n[1]; M[4] += n[1] * n[2]; M[5] += n[2] * n[2];
            r[0] += n[0] * d; r[1] += n[1] * d; r[2] += n[2] * d;
            for (int a = 0; a < 3; ++a)
                if (fabs(p[a]) > 1e-8) { msum[a] += p[a]; ++mcnt[a]; }      // np.isclose(x, 0) coordinates are ignored
        }
        has_v[v] = any ? 1 : 0;
        if (!any) continue;
        const double b2 = bias * bias;
        for (int a = 0; a < 3; ++a) {
            const double mass = msum[a] / (double)mcnt[a];       // 0/0 -> NaN, like nanmean of an empty slice
            r[a] += b2 * mass;
        }
        M[0] += b2; M[3] += b2; M[5] += b2;
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

// lattice index of a voxel centre; code on a lattice padded by one cell per axis (neighbour codes never alias)
struct Lattice {
    double x0, y0, z0;      // centre of lattice cell (0, 0, 0) (shift included)
    double dx, dy, dz;
    long long ny1, nz1;     // padded extents
};
__device__ __forceinline__ void lattice_ijk(const Lattice& L, double x, double y, double z, long long& i, long long& j, long long& k) {
    i = llrint((x - L.x0) / L.dx);
    j = llrint((y - L.y0) / L.dy);
    k = llrint((z - L.z0) / L.dz);
}
__device__ __forceinline__ unsigned long long lattice_code(const Lattice& L, long long i, long long j, long long k) {
    return (unsigned long long)((i * L.ny1 + j) * L.nz1 + k);
}

constexpr unsigned long long kEmpty = ~0ull;
__device__ __forceinline__ unsigned long long hash64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// compact the vertices (voxel order) and insert (lattice code -> vertex id) into the hash table
__global__ void compact_vertices_kernel(const unsigned char* __restrict__ has_v, long long nvox, const long long* __restrict__ offsets,
                                        const double* __restrict__ vert, const double* __restrict__ cen, long long ld_cen, Lattice L,
                                        int* __restrict__ vid, double* __restrict__ vert_c, long long ld_v,
                                        unsigned long long* __restrict__ keys, int* __restrict__ vals, unsigned long long cap_mask) {
    __shared__ int warp_off[kScanB / 32];
    const long long v = (long long)blockIdx.x * kScanB + threadIdx.x;
    const int m = (v < nvox) ? (int)has_v[v] : 0;
    const long long p = block_rank(m, offsets, warp_off);
    if (v < nvox) vid[v] = m ? (int)p : -1;
    if (m) {
        vert_c[p] = vert[v];
        vert_c[ld_v + p] = vert[nvox + v];
        vert_c[2 * ld_v + p] = vert[2 * nvox + v];
        long long i, j, k;
        lattice_ijk(L, cen[v], cen[ld_cen + v], cen[2 * ld_cen + v], i, j, k);
        const unsigned long long code = lattice_code(L, i, j, k);
        unsigned long long h = hash64(code) & cap_mask;
        while (true) {
            const unsigned long long old = atomicCAS(&keys[h], kEmpty, code);
            if (old == kEmpty || old == code) { vals[h] = (int)p; break; }
            h = (h + 1) & cap_mask;
        }
    }
}

__device__ __forceinline__ int hash_find(const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                         unsigned long long cap_mask, unsigned long long code) {
    unsigned long long h = hash64(code) & cap_mask;
    while (true) {
        const unsigned long long kk = keys[h];
        if (kk == code) return vals[h];
        if (kk == kEmpty) return -1;
        h = (h + 1) & cap_mask;
    }
}

// triangle candidates: slot t = ax * nvox + v; edge hh[ax] of voxel v crossed and the three neighbours across it exist
__device__ __forceinline__ bool tri_corners(int ax, long long v, const unsigned char* __restrict__ valid, const double* __restrict__ cen,
                                            long long ld_cen, const Lattice& L, const unsigned long long* __restrict__ keys,
                                            const int* __restrict__ vals, unsigned long long cap_mask, int& a, int& b, int& c) {
    const int hh = (ax == 0) ? 3 : (ax == 1 ? 7 : 11);
    if (!valid[12 * v + hh]) return false;
    long long i, j, k;
    lattice_ijk(L, cen[v], cen[ld_cen + v], cen[2 * ld_cen + v], i, j, k);
    const int u = (ax == 0) ? 1 : 0, w = (ax == 2) ? 1 : 2;       // the two other axes, in increasing order
    long long ku[3] = {i, j, k}, kv[3] = {i, j, k}, kuv[3] = {i, j, k};
    ku[u] += 1; kv[w] += 1; kuv[u] += 1; kuv[w] += 1;
    a = hash_find(keys, vals, cap_mask, lattice_code(L, ku[0], ku[1], ku[2]));
    b = hash_find(keys, vals, cap_mask, lattice_code(L, kv[0], kv[1], kv[2]));
    c = hash_find(keys, vals, cap_mask, lattice_code(L, kuv[0], kuv[1], kuv[2]));
    return a >= 0 && b >= 0 && c >= 0;
}

__global__ void tri_flags_kernel(const unsigned char* __restrict__ valid, long long nvox, const double* __restrict__ cen, long long ld_cen,
                                 Lattice L, const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                 unsigned long long cap_mask, unsigned char* __restrict__ tflag) {
    const long long total = 3 * nvox;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ax = (int)(t / nvox);
        const long long v = t - (long long)ax * nvox;
        int a, b, c;
        tflag[t] = tri_corners(ax, v, valid, cen, ld_cen, L, keys, vals, cap_mask, a, b, c) ? 1 : 0;
    }
}

__global__ void tri_emit_kernel(const unsigned char* __restrict__ tflag, const unsigned char* __restrict__ valid, long long nvox,
                                const long long* __restrict__ offsets, const double* __restrict__ cen, long long ld_cen, Lattice L,
                                const unsigned long long* __restrict__ keys, const int* __restrict__ vals, unsigned long long cap_mask,
                                const int* __restrict__ vid, int* __restrict__ tris) {
    __shared__ int warp_off[kScanB / 32];
    const long long total = 3 * nvox;
    const long long t = (long long)blockIdx.x * kScanB + threadIdx.x;
    const int m = (t < total) ? (int)tflag[t] : 0;
    const long long p = block_rank(m, offsets, warp_off);
    if (m) {
        const int ax = (int)(t / nvox);
        const long long v = t - (long long)ax * nvox;
        int a, b, c;
        tri_corners(ax, v, valid, cen, ld_cen, L, keys, vals, cap_mask, a, b, c);
        const int n0 = vid[v];
        int* o = tris + 6 * p;
        o[0] = n0; o[1] = a; o[2] = c;
        o[3] = n0; o[4] = c; o[5] = b;
    }
}

__global__ void fill_u64_kernel(unsigned long long* p, long long n, unsigned long long v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void store_counts_kernel(const long long* n_valid, const long long* n_vert, const long long* n_tri, long long* counts) {
    counts[0] = *n_valid;
    counts[1] = *n_vert;
    counts[2] = 2 * *n_tri;
}


// ---- corner de-duplication of an octree level --------------------------------------------------------------------------
// The 8 corners of neighbouring voxels coincide (a lattice corner belongs to up to 8 voxels; among the children of one
// parent 27 of the 64 corner slots are distinct).  The reference evaluates every slot; here a hash table over the corner
// lattice finds, for every slot, the FIRST slot with the same lattice corner (atomicMin on the slot index: deterministic),
// the owners are compacted into the list of unique corners, and the fields are evaluated on that list only; the level's
// corner segment is then filled by a gather (gpb_expand_rows).  Duplicates therefore carry exactly the owner's value.
__device__ __forceinline__ unsigned long long corner_code(const Lattice& L, const double* __restrict__ cen, long long ld_c, long long e) {
    const long long v = e >> 3;
    const int c = (int)(e & 7);
    long long i, j, k;
    lattice_ijk(L, cen[v], cen[ld_c + v], cen[2 * ld_c + v], i, j, k);
    return lattice_code(L, i + ((c >> 2) & 1), j + ((c >> 1) & 1), k + (c & 1));
}

__global__ void corner_insert_kernel(const double* __restrict__ cen, long long ld_c, long long nslots, Lattice L,
                                     unsigned long long* __restrict__ keys, int* __restrict__ vals, unsigned long long cap_mask) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nslots; e += (long long)gridDim.x * blockDim.x) {
        const unsigned long long code = corner_code(L, cen, ld_c, e);
        unsigned long long h = hash64(code) & cap_mask;
        while (true) {
            const unsigned long long old = atomicCAS(&keys[h], kEmpty, code);
            if (old == kEmpty || old == code) { atomicMin(&vals[h], (int)e); break; }
            h = (h + 1) & cap_mask;
        }
    }
}

__global__ void corner_owner_kernel(const double* __restrict__ cen, long long ld_c, long long nslots, Lattice L,
                                    const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                    unsigned long long cap_mask, int* __restrict__ owner, unsigned char* __restrict__ is_owner) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nslots; e += (long long)gridDim.x * blockDim.x) {
        const int o = hash_find(keys, vals, cap_mask, corner_code(L, cen, ld_c, e));
        owner[e] = o;
        is_owner[e] = (o == (int)e) ? 1 : 0;
    }
}

__global__ void fill_int_kernel(int* p, long long n, int v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// unique corner coordinates (same expression as corners_kernel: centre +- half cell) and the rank of every owner
__global__ void corner_emit_kernel(const unsigned char* __restrict__ is_owner, long long nslots, const long long* __restrict__ offsets,
                                   const double* __restrict__ cen, long long ld_c, double hx, double hy, double hz,
                                   int* __restrict__ uid, double* __restrict__ xyz_u, long long ld_u) {
    __shared__ int warp_off[kScanB / 32];
    const long long e = (long long)blockIdx.x * kScanB + threadIdx.x;
    const int m = (e < nslots) ? (int)is_owner[e] : 0;
    const long long p = block_rank(m, offsets, warp_off);
    if (e < nslots) uid[e] = m ? (int)p : -1;
    if (m) {
        const long long v = e >> 3;
        const int c = (int)(e & 7);
        xyz_u[p] = cen[v] + ((c & 4) ? hx : -hx);
        xyz_u[ld_u + p] = cen[ld_c + v] + ((c & 2) ? hy : -hy);
        xyz_u[2 * ld_u + p] = cen[2 * ld_c + v] + ((c & 1) ? hz : -hz);
    }
}

__global__ void corner_map_kernel(const int* __restrict__ owner, const int* __restrict__ uid, long long nslots, int* __restrict__ map) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nslots; e += (long long)gridDim.x * blockDim.x)
        map[e] = uid[owner[e]];
}

__global__ void add_base_kernel(const long long* n, long long base, long long* out) { *out = base + *n; }

__global__ void expand_rows_kernel(const double* __restrict__ src, long long ld_src, const int* __restrict__ map, int n_rows,
                                   long long count, double* __restrict__ dst, long long ld_dst) {
    const long long total = (long long)n_rows * count;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / count, e = t - r * count;
        dst[r * ld_dst + e] = src[r * ld_src + map[e]];
    }
}

}  // namespace

extern "C" long long gpb_dc_scratch_bytes(long long nvox) {
    if (nvox <= 0) return 256;
    const long long n12 = 12 * nvox, n3 = 3 * nvox;
    const long long b12 = (n12 + kScanB - 1) / kScanB + 1, b1 = (nvox + kScanB - 1) / kScanB + 1, b3 = (n3 + kScanB - 1) / kScanB + 1;
    long long cap = 16;
    while (cap < 2 * nvox) cap <<= 1;
    long long bytes = 0;
    auto add = [&](long long b) { bytes += (b + 255) / 256 * 256; };
    add(3 * n12 * 8);        // xyz_e (uncompacted crossings)
    add(n12 * 4);            // pos
    add(b12 * 8); add(b1 * 8); add(b3 * 8);
    add(3 * nvox * 8);       // vert (per voxel)
    add(nvox);               // has_v
    add(nvox * 4);           // vid
    add(cap * 8); add(cap * 4);
    add(n3);                 // tflag
    add(3 * n12 * 8);        // Z scratch of the gradient evaluation is not needed; gradient of the crossings when the caller keeps none
    return bytes;
}

// One surface.  corners [3][ld_k] (8 per voxel), Z_corners [8 nvox] of the stack's field, sq_corners [8 nvox] squeezed mask of
// the stack at the corners (NULL for fault stacks: every voxel), centers [3][ld_c]; lattice = the level's voxel lattice
// (cell (0,0,0) centre incl. shift, cell size, number of cells).  Outputs (caller-allocated for the worst case):
// valid [12 nvox], xyz_c / grad_c [3][12 nvox] (compacted crossings and their gradients; grad_c may be NULL),
// vertices [3][nvox] (compacted), triangles [6 nvox][3] int32, counts [3] (crossings, vertices, triangles; device).
extern "C" int gpb_dual_contour(const gpb_stack* st, const double* eval_table, const double* corners, long long ld_k,
                                const double* Z_corners, const unsigned char* sq_corners, const double* centers, long long ld_c,
                                long long nvox, const double* iso_dev, const gpb_regular_grid* lattice, double bias,
                                void* scratch, long long scratch_bytes, unsigned char* valid, double* xyz_c, double* grad_c,
                                double* vertices, int* triangles, long long* counts, void* stream) {
    GPB_REQUIRE(st && eval_table && corners && Z_corners && centers && iso_dev && lattice && valid && xyz_c && vertices && triangles && counts,
                "null argument");
    GPB_REQUIRE(nvox >= 0 && ld_k >= 8 * nvox && ld_c >= nvox, "bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    if (nvox == 0) {
        GPB_CHECK_CUDA(cudaMemsetAsync(counts, 0, 3 * sizeof(long long), s));
        return GPB_OK;
    }
    GPB_REQUIRE(scratch && scratch_bytes >= gpb_dc_scratch_bytes(nvox), "scratch too small (gpb_dc_scratch_bytes)");
    const long long n12 = 12 * nvox, n3 = 3 * nvox;
    const long long nb12 = (n12 + kScanB - 1) / kScanB, nb1 = (nvox + kScanB - 1) / kScanB, nb3 = (n3 + kScanB - 1) / kScanB;
    unsigned long long cap = 16;
    while (cap < (unsigned long long)(2 * nvox)) cap <<= 1;
    char* p = (char*)scratch;
    auto take = [&](long long b) { char* q = p; p += (b + 255) / 256 * 256; return q; };
    double* xyz_e = (double*)take(3 * n12 * 8);
    int* pos = (int*)take(n12 * 4);
    long long* off12 = (long long*)take((nb12 + 1) * 8);
    long long* off1 = (long long*)take((nb1 + 1) * 8);
    long long* off3 = (long long*)take((nb3 + 1) * 8);
    double* vert = (double*)take(3 * nvox * 8);
    unsigned char* has_v = (unsigned char*)take(nvox);
    int* vid = (int*)take(nvox * 4);
    unsigned long long* keys = (unsigned long long*)take(cap * 8);
    int* vals = (int*)take(cap * 4);
    unsigned char* tflag = (unsigned char*)take(n3);
    double* grad_own = (double*)take(3 * n12 * 8);
    double* grad = grad_c ? grad_c : grad_own;

    Lattice L;
    L.x0 = lattice->x0; L.y0 = lattice->y0; L.z0 = lattice->z0;
    L.dx = lattice->dx; L.dy = lattice->dy; L.dz = lattice->dz;
    L.ny1 = (long long)lattice->ny + 2; L.nz1 = (long long)lattice->nz + 2;

    // 1. crossings
    dc_edges2_kernel<<<grid_for(n12), kT, 0, s>>>(corners, ld_k, Z_corners, nvox, iso_dev, sq_corners, valid, xyz_e);
    GPB_LAUNCH_CHECK();
    // 2. compaction
    block_count_kernel<<<(unsigned)nb12, kScanB, 0, s>>>(valid, n12, off12);
    GPB_LAUNCH_CHECK();
    scan_blocks_kernel<<<1, 1024, 0, s>>>(off12, nb12, off12 + nb12);
    GPB_LAUNCH_CHECK();
    compact_edges_kernel<<<(unsigned)nb12, kScanB, 0, s>>>(valid, n12, off12, xyz_e, pos, xyz_c, n12);
    GPB_LAUNCH_CHECK();
    // 3. gradient of the stack's field at the crossings (the fault drift has no gradient term: fault columns skipped);
    //    the scalar field itself is written over the now unused uncompacted crossing buffer
    {
        gpb_stack s0 = *st;
        s0.n_faults = 0;
        GpbEvalCall c;
        c.st = &s0;
        c.src = eval_table;
        c.xyz = xyz_c;
        c.ld_xyz = n12;
        c.m = n12;
        c.m_dev = off12 + nb12;
        c.Z = xyz_e;
        c.gx = grad; c.gy = grad + n12; c.gz = grad + 2 * n12;
        int rc = gpb_eval_call(c, s);
        if (rc) return rc;
    }
    // 4. vertices
    dc_vertices2_kernel<<<grid_for(nvox), kT, 0, s>>>(valid, pos, xyz_c, grad, n12, nvox, bias, vert, has_v);
    GPB_LAUNCH_CHECK();
    block_count_kernel<<<(unsigned)nb1, kScanB, 0, s>>>(has_v, nvox, off1);
    GPB_LAUNCH_CHECK();
    scan_blocks_kernel<<<1, 1024, 0, s>>>(off1, nb1, off1 + nb1);
    GPB_LAUNCH_CHECK();
    fill_u64_kernel<<<grid_for((long long)cap), kT, 0, s>>>(keys, (long long)cap, kEmpty);
    GPB_LAUNCH_CHECK();
    compact_vertices_kernel<<<(unsigned)nb1, kScanB, 0, s>>>(has_v, nvox, off1, vert, centers, ld_c, L, vid, vertices, nvox, keys, vals, cap - 1);
    GPB_LAUNCH_CHECK();
    // 5. triangles
    tri_flags_kernel<<<grid_for(n3), kT, 0, s>>>(valid, nvox, centers, ld_c, L, keys, vals, cap - 1, tflag);
    GPB_LAUNCH_CHECK();
    block_count_kernel<<<(unsigned)nb3, kScanB, 0, s>>>(tflag, n3, off3);
    GPB_LAUNCH_CHECK();
    scan_blocks_kernel<<<1, 1024, 0, s>>>(off3, nb3, off3 + nb3);
    GPB_LAUNCH_CHECK();
    tri_emit_kernel<<<(unsigned)nb3, kScanB, 0, s>>>(tflag, valid, nvox, off3, centers, ld_c, L, keys, vals, cap - 1, vid, triangles);
    GPB_LAUNCH_CHECK();
    store_counts_kernel<<<1, 1, 0, s>>>(off12 + nb12, off1 + nb1, off3 + nb3, counts);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}


// ---- corner de-duplication (see the kernels above) -----------------------------------------------------------------------
extern "C" long long gpb_corner_scratch_bytes(long long nvox) {
    if (nvox <= 0) return 256;
    const long long ns = 8 * nvox;
    unsigned long long cap = 16;
    while (cap < (unsigned long long)(2 * ns)) cap <<= 1;
    const long long nb = (ns + kScanB - 1) / kScanB + 1;
    long long bytes = 0;
    auto add = [&](long long b) { bytes += (b + 255) / 256 * 256; };
    add((long long)cap * 8); add((long long)cap * 4); add(ns * 4); add(ns); add(nb * 8); add(ns * 4);
    return bytes;
}

// Step 1: classify the 8 nvox corner slots (owner = first slot on the same lattice corner) and count the unique corners
// (returned on the host: synchronises).  Step 2 (gpb_corner_unique_emit, same scratch): unique coordinates [3][ld_u] and the
// slot -> unique index map [8 nvox].
extern "C" int gpb_corner_unique_count(const double* centers, long long ld_c, long long nvox, const gpb_regular_grid* lattice,
                                       void* scratch, long long scratch_bytes, long long* n_unique_host, void* stream) {
    GPB_REQUIRE(centers && lattice && nvox >= 0 && ld_c >= nvox, "bad arguments");
    if (n_unique_host) *n_unique_host = 0;
    if (nvox == 0) return GPB_OK;
    GPB_REQUIRE(8 * nvox < 2000000000LL, "too many corner slots for 32-bit slot indices");
    GPB_REQUIRE(scratch && scratch_bytes >= gpb_corner_scratch_bytes(nvox), "scratch too small (gpb_corner_scratch_bytes)");
    cudaStream_t s = (cudaStream_t)stream;
    const long long ns = 8 * nvox;
    unsigned long long cap = 16;
    while (cap < (unsigned long long)(2 * ns)) cap <<= 1;
    const long long nb = (ns + kScanB - 1) / kScanB;
    char* p = (char*)scratch;
    auto take = [&](long long b) { char* q = p; p += (b + 255) / 256 * 256; return q; };
    unsigned long long* keys = (unsigned long long*)take((long long)cap * 8);
    int* vals = (int*)take((long long)cap * 4);
    int* owner = (int*)take(ns * 4);
    unsigned char* is_owner = (unsigned char*)take(ns);
    long long* off = (long long*)take((nb + 1) * 8);
    Lattice L;
    L.x0 = lattice->x0; L.y0 = lattice->y0; L.z0 = lattice->z0;
    L.dx = lattice->dx; L.dy = lattice->dy; L.dz = lattice->dz;
    L.ny1 = (long long)lattice->ny + 2; L.nz1 = (long long)lattice->nz + 2;
    fill_u64_kernel<<<grid_for((long long)cap), kT, 0, s>>>(keys, (long long)cap, kEmpty);
    GPB_LAUNCH_CHECK();
    fill_int_kernel<<<grid_for((long long)cap), kT, 0, s>>>(vals, (long long)cap, 0x7fffffff);
    GPB_LAUNCH_CHECK();
    corner_insert_kernel<<<grid_for(ns), kT, 0, s>>>(centers, ld_c, ns, L, keys, vals, cap - 1);
    GPB_LAUNCH_CHECK();
    corner_owner_kernel<<<grid_for(ns), kT, 0, s>>>(centers, ld_c, ns, L, keys, vals, cap - 1, owner, is_owner);
    GPB_LAUNCH_CHECK();
    block_count_kernel<<<(unsigned)nb, kScanB, 0, s>>>(is_owner, ns, off);
    GPB_LAUNCH_CHECK();
    scan_blocks_kernel<<<1, 1024, 0, s>>>(off, nb, off + nb);
    GPB_LAUNCH_CHECK();
    if (n_unique_host != nullptr) {
        GPB_CHECK_CUDA(cudaMemcpyAsync(n_unique_host, off + nb, sizeof(long long), cudaMemcpyDeviceToHost, s));
        GPB_CHECK_CUDA(cudaStreamSynchronize(s));
    }
    return GPB_OK;
}

extern "C" int gpb_corner_unique_emit(const double* centers, long long ld_c, long long nvox, double hx, double hy, double hz,
                                      void* scratch, long long scratch_bytes, double* xyz_unique, long long ld_u, int* map,
                                      long long count_base, long long* count_dev_out, void* stream) {
    GPB_REQUIRE(centers && xyz_unique && map && nvox >= 0 && ld_c >= nvox, "bad arguments");
    if (nvox == 0) return GPB_OK;
    GPB_REQUIRE(scratch && scratch_bytes >= gpb_corner_scratch_bytes(nvox), "scratch too small (gpb_corner_scratch_bytes)");
    cudaStream_t s = (cudaStream_t)stream;
    const long long ns = 8 * nvox;
    unsigned long long cap = 16;
    while (cap < (unsigned long long)(2 * ns)) cap <<= 1;
    const long long nb = (ns + kScanB - 1) / kScanB;
    char* p = (char*)scratch;
    auto take = [&](long long b) { char* q = p; p += (b + 255) / 256 * 256; return q; };
    take((long long)cap * 8);
    take((long long)cap * 4);
    int* owner = (int*)take(ns * 4);
    unsigned char* is_owner = (unsigned char*)take(ns);
    long long* off = (long long*)take((nb + 1) * 8);
    int* uid = (int*)take(ns * 4);
    corner_emit_kernel<<<(unsigned)nb, kScanB, 0, s>>>(is_owner, ns, off, centers, ld_c, hx, hy, hz, uid, xyz_unique, ld_u);
    GPB_LAUNCH_CHECK();
    corner_map_kernel<<<grid_for(ns), kT, 0, s>>>(owner, uid, ns, map);
    GPB_LAUNCH_CHECK();
    if (count_dev_out != nullptr) {
        add_base_kernel<<<1, 1, 0, s>>>(off + nb, count_base, count_dev_out);
        GPB_LAUNCH_CHECK();
    }
    return GPB_OK;
}

// dst[r][e] = src[r][map[e]], r < n_rows, e < count (fills a level's corner segment from the unique-corner results)
extern "C" int gpb_expand_rows(const double* src, long long ld_src, const int* map, int n_rows, long long count, double* dst,
                               long long ld_dst, void* stream) {
    GPB_REQUIRE(src && map && dst && n_rows >= 0 && count >= 0, "bad arguments");
    if (n_rows == 0 || count == 0) return GPB_OK;
    expand_rows_kernel<<<grid_for((long long)n_rows * count), kT, 0, (cudaStream_t)stream>>>(src, ld_src, map, n_rows, count, dst, ld_dst);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}
