/* gempy_b200.h -- C ABI of the B200 backend for GemPy's implicit co-kriging hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every `double*` / `int*`
 * below is a DEVICE pointer unless the parameter is documented as host; `stream` is a cudaStream_t
 * passed as void* (0 = legacy default stream).  Every entry returns 0 on success or a negative
 * GPB_E_* code, in which case gpb_last_error() holds a message.  Launches are asynchronous on
 * `stream` unless stated otherwise.  Entries act on the CURRENT CUDA device (cudaSetDevice before calling; the host
 * mirror does).  Process-wide state: the launch counter (atomic); the tuning setters gpb_lu_set_outer_*; per DEVICE:
 * kernel attributes, one high-priority side stream + two events used by the factorisations' look-ahead, and one
 * stream-ordered memory pool for the entries' own scratch (flags, scan counts, inverse diagonal blocks: a few MB; the pool
 * keeps what it has allocated until the process ends, GPB_SCRATCH_POOL=0 reverts to the device's default pool) -- the
 * enqueue of a factorisation is serialised per device by a mutex inside the library, so calls are thread-safe, but a
 * factor and the gpb_lu_apply calls that replay it must see the same gpb_lu_set_outer_* settings.
 *
 * Reference interface replaced (the engine package itself is not vendored in the reference tree; the
 * only reference-side binding is the Python call
 *     gempy_engine.compute_model(interpolation_input, options, data_descriptor, geophysics_input)
 * at /root/reference/gempy/API/compute_API.py:68-73, :140-145 and
 * gempy/modules/optimize_nuggets/_ops.py:18-23).  gempy_b200/engine/compute.py is the host mirror of
 * that call; it binds the entry points below with ctypes (gempy_b200/_lib.py).  Per-entry notes say
 * which engine stage each one implements.
 *
 * Layout conventions
 *  - every coordinate table is SoA: [3][n] doubles (x row, y row, z row), in the TRANSFORMED
 *    coordinate system the bridge produces (_engine_factory.py:26-37);
 *  - system rows/cols are ordered [G_x(n_ori) | G_y | G_z | increments (n_rest) | drift | faults];
 *  - matrices are column-major with leading dimension lda (they are symmetric on assembly).
 */
#ifndef GEMPY_B200_H
#define GEMPY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPB_OK 0
#define GPB_E_INVALID  (-1)   /* bad argument */
#define GPB_E_CUDA     (-2)   /* CUDA runtime error */
#define GPB_E_SINGULAR (-3)   /* zero pivot in the LU */
#define GPB_E_NODEVICE (-4)   /* no usable sm_100 device */

#define GPB_KERNEL_CUBIC       0
#define GPB_KERNEL_EXPONENTIAL 1   /* exp(-r^2 / (2 a^2)) */
#define GPB_KERNEL_MATERN52    2

#define GPB_REL_ERODE    1   /* StackRelationType encodings of the reference's serialization goldens */
#define GPB_REL_ONLAP    2
#define GPB_REL_FAULT    3
#define GPB_REL_BASEMENT 4

/* One stack (= one scalar field) after the ref/rest split.
 * Mirrors what the engine's preprocess stage produces from SurfacePoints / Orientations
 * (_engine_factory.py:27-37) for one structural group (structural_frame.py:333-350). */
typedef struct gpb_stack {
    int n_ori;              /* orientations */
    int n_rest;             /* increments rest_i - ref_i  (= n_sp - n_surfaces) */
    int n_surf;             /* surfaces in the stack */
    int n_drift;            /* 0, 3 (degree 1) or 9 (degree 2) */
    int n_faults;           /* fault-drift columns */
    int kernel;             /* GPB_KERNEL_* */
    double range, c_o, i_res, gi_res;
    const double* ori_pos;      /* [3][n_ori] */
    const double* ori_grad;     /* [3][n_ori] */
    const double* ori_nugget;   /* [n_ori]    */
    const double* rest;         /* [3][n_rest] */
    const double* ref;          /* [3][n_rest]  reference point of each increment's surface */
    const double* sp_nugget;    /* [n_rest]  row nugget of each increment */
    const double* fault_rest;   /* [n_faults][n_rest] fault-block value at rest_i (may be NULL if n_faults == 0) */
    const double* fault_ref;    /* [n_faults][n_rest] */
    const int*    surf_offsets; /* [n_surf+1] first increment row of every surface */
    const double* ref_unique;   /* [3][n_surf] the reference point of each surface */
} gpb_stack;

/* Regular (dense or octree-root) grid: cell centres, x slowest / z fastest
 * (gempy/core/data/grid_modules/regular_grid.py:58-71), every centre displaced by `shift`. */
typedef struct gpb_regular_grid {
    double x0, y0, z0;      /* centre of cell (0,0,0), shift included */
    double dx, dy, dz;
    int nx, ny, nz;
} gpb_regular_grid;

/* ---- library ---------------------------------------------------------------------------------- */
const char* gpb_last_error(void);
int  gpb_version(void);
/* Number of SMs / compute capability of `device`; GPB_E_NODEVICE if it is not sm_100. (host) */
int  gpb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);
/* Counter of kernels this library has launched since load (host; used by bench.py's gpu_launches). */
long long gpb_launch_count(void);

/* Measured FP64 FMA throughput of the current device (register-resident DFMA chains, 8 per thread), in
 * TFLOP/s (FMA = 2).  Used as the roofline denominator of the evaluation kernel.  Synchronises. (host out) */
int gpb_bench_dfma(int iters, double* tflops_host, void* stream);
/* Same for the FP64 tensor-core path (mma.sync m8n8k4 chains): decides whether the LU trailing update belongs on
 * DMMA or on the DFMA pipe on this part. */
int gpb_bench_dmma(int iters, double* tflops_host, void* stream);
/* Both at once: every warp issues `ratio` (4, 8, 16 or 32) DFMA per DMMA.  Tells whether the tensor sub-pipe and the
 * FP64 FMA pipe run concurrently. */
int gpb_bench_mixed(int iters, int ratio, double* dfma_tflops_host, double* dmma_tflops_host, void* stream);

/* ---- (1) covariance assembly  [engine stage "kernel_constructor", SURVEY 8a2 row (1)] ------------ */
/* n = 3*n_ori + n_rest + n_drift + n_faults.  Writes the full symmetric n x n matrix A (lda >= n) and the
 * right-hand side b = [G_x; G_y; G_z; 0]. */
int gpb_system_size(const gpb_stack* st);
int gpb_assemble_cov(const gpb_stack* st, double* A, int lda, double* b, void* stream);
/* flags: GPB_COV_LOWER_ONLY writes the lower triangle (+ diagonal) only -- all the symmetric solve reads; for systems of
 * order >= 512 this halves the HBM traffic of the assembly (smaller systems are always written in full). */
#define GPB_COV_LOWER_ONLY 1
int gpb_assemble_cov_ex(const gpb_stack* st, double* A, int lda, double* b, int flags, void* stream);

/* ---- (2) dense solve  [engine stage "solver", kernel_solver = 1 (direct)] ------------------------- */
/* In-place blocked right-looking LU with partial pivoting (DMMA trailing updates), then the triangular
 * solves.  A is overwritten with L\U, b (n x nrhs, ldb) with the solution, ipiv (n ints) with the pivot rows.
 * `info` (device int, may be NULL) receives 0 or the 1-based index of a zero pivot. */
int gpb_lu_solve(int n, double* A, int lda, double* b, int nrhs, int ldb, int* ipiv, int* info, void* stream);
/* Systems with n >= min_n are factored in outer blocks of 256 columns (one K = 256 trailing update per block instead
 * of eight K = 32 ones: 8x less HBM traffic, which is what bounds the solve for n in the tens of thousands).  Default
 * 10240 (or the environment variable GPB_LU_OUTER_MIN_N).  Returns the previous value; min_n < 0 only queries.
 * A factorisation and the gpb_lu_apply calls that use it must run under the same setting. (host) */
int gpb_lu_set_outer_min_n(int min_n);
/* Outer block width of that schedule: 0 = default (256), or 128 / 256 to force it (tests, tuning); other values are
 * ignored.  Returns the previous setting. (host) */
int gpb_lu_set_outer_width(int width);
/* Factor only / solve only (weight reuse across octree levels, timing against cusolverDnDgetrf/Dgetrs). */
int gpb_lu_factor(int n, double* A, int lda, int* ipiv, int* info, void* stream);
int gpb_lu_apply(int n, const double* LU, int lda, const int* ipiv, double* b, int nrhs, int ldb, void* stream);

/* Symmetric path for the saddle-point system A = [K U; U^T 0] the covariance assembly produces: K = leading nk x nk
 * block (symmetric positive definite: covariances + nuggets), U = the n - nk <= 64 universal-drift and fault-drift
 * columns.  Blocked right-looking Cholesky of K on the LOWER triangle (no pivoting, DMMA trailing updates), Schur
 * complement for the drift coefficients, triangular solves.  Only the lower triangle of A is read; A is overwritten.
 * The right-hand sides ride through the factorisation as extra rows: lda >= n + nrhs (rows n .. n+nrhs-1 of every
 * column are workspace).  b (n x nrhs, ldb) is overwritten with the solution.  `info` (device int, may be NULL): 0, or
 * the 1-based column of a non-positive pivot -- K is then not numerically positive definite (or U is rank deficient)
 * and the caller should re-assemble and use gpb_lu_solve. */
int gpb_sym_solve(int n, int nk, double* A, int lda, double* b, int nrhs, int ldb, int* info, void* stream);

/* ---- (3) fused field + gradient evaluation  [engine stage "evaluator"; the dominant kernel] -------- */
/* Pack the weights of one solved stack into the evaluation source tables (range-normalised coordinates,
 * pre-scaled weights, one aggregated source per reference point).  `src` must hold gpb_eval_table_doubles()
 * doubles. */
long long gpb_eval_table_doubles(const gpb_stack* st);
int gpb_pack_eval_table(const gpb_stack* st, const double* w, double* src, void* stream);
/* Evaluate Z (and, if gx/gy/gz != NULL, the engine-convention gradient) at grid points [i0, i1) of a
 * regular grid; outputs are indexed from 0 (out[k] is point i0 + k).  fault_vals: [n_faults][ld_fault] values
 * of the active fault blocks at the same points (NULL if n_faults == 0). */
int gpb_eval_regular(const gpb_stack* st, const double* src, const gpb_regular_grid* grid,
                     long long i0, long long i1, const double* fault_vals, long long ld_fault,
                     double* Z, double* gx, double* gy, double* gz, void* stream);
/* Same at an explicit point list xyz = [3][ld_xyz] (octree levels, corners, custom grids, surface points). */
int gpb_eval_points(const gpb_stack* st, const double* src, const double* xyz, long long ld_xyz, long long m,
                    const double* fault_vals, long long ld_fault,
                    double* Z, double* gx, double* gy, double* gz, void* stream);

/* ---- (4a) activator + masks + combination  [engine stages "activator", mask/combine] -------------- */
/* block[k] = sum_j ids[j] * (sigma(slope (Z - iso[j])) - sigma(slope (Z - iso[j-1])));  ids has n_surf+1 entries. */
int gpb_activate(const double* Z, long long m, const double* isovalues, const double* ids, int n_surf,
                 double slope, double* block, void* stream);
/* Min over a device array (fault blocks are shifted by their minimum before they act as drift). */
int gpb_min(const double* v, long long m, double* out_min, void* stream);
int gpb_shift(const double* v, long long m, const double* minus, double* out, void* stream);
/* Combine stacks top-down.  Z/block: [n_stacks][ld]; relations: host int[n_stacks]; iso_min/iso_max: device
 * [n_stacks] (min / max isovalue of each stack).  Outputs: final_block[m], faults_block[m],
 * squeezed_mask[n_stacks][ld] (uint8), mask[n_stacks][ld] (uint8, may be NULL). */
int gpb_combine(const double* Z, const double* block, long long ld, long long m, int n_stacks,
                const int* relations_host, const double* iso_min, const double* iso_max,
                double* final_block, double* faults_block, unsigned char* squeezed_mask, unsigned char* mask,
                void* stream);

/* ---- level executor: every stack of a model on one evaluation domain  [engine stage interpolate_all_fields] ------
 * One C call per octree level replaces the per-stack Python loop of the reference (and of round 1 of this backend):
 * solve (level 0), fused evaluation + fault drift + activator per stack, combination.  All device buffers belong to the
 * caller; the handle stores the description.  Stacks are processed in order; a stack's fault drift comes from EARLIER
 * fault stacks (reference: structural_frame.py:243-273 fault_relations, upper triangular). */
typedef struct gpb_model_stack {
    gpb_stack st;                  /* tables of the stack; st.fault_rest / st.fault_ref: caller-allocated
                                      [n_faults][n_rest] buffers FILLED by gpb_model_solve_stack */
    int relation;                  /* GPB_REL_* */
    const int* fault_stacks_host;  /* host [st.n_faults]: indices of the fault stacks drifting this one (read at create) */
    const int* fault_stacks_dev;   /* the same list on the device */
    int sp_begin;                  /* first surface point of the stack in the model-wide table */
    int n_sp;                      /* its surface points (n_rest + n_surf) */
    const double* unit_ids;        /* device [n_surf + 1] unit values of the surfaces + the unit below */
    double* weights;               /* device [n] out */
    double* eval_table;            /* device [gpb_eval_table_doubles()] out */
    double* isovalues;             /* device [n_surf] out: scalar field at each surface's reference point */
} gpb_model_stack;

typedef struct gpb_model_desc {
    int n_stacks;                  /* <= 64 */
    const gpb_model_stack* stacks; /* host [n_stacks] */
    const double* sp_all;          /* device [3][n_sp_all]: every surface point of the model, stack after stack */
    long long n_sp_all;
    double sigmoid_slope;
    double* iso_min;               /* device [n_stacks] out */
    double* iso_max;               /* device [n_stacks] out */
    double* fault_min;             /* device [n_stacks] out: minimum of each fault stack's block on the current level */
    int solver;                    /* 0: symmetric path, pivoted LU for n <= 160 or when not positive definite; 1: LU only */
} gpb_model_desc;

#define GPB_SEG_POINTS  0
#define GPB_SEG_REGULAR 1
#define GPB_SEG_OCTETS  2      /* POINTS whose entries 8g .. 8g+7 are the 8 children of one voxel (what gpb_emit_marked writes):
                                  evaluated by the octet kernel, which shares the per-axis distances among the 8 points */
typedef struct gpb_segment {
    int kind;
    long long count;               /* points of the segment */
    long long out_offset;          /* their position in the level's outputs */
    const double* xyz;             /* POINTS: device [3][ld_xyz] */
    long long ld_xyz;
    gpb_regular_grid grid;         /* REGULAR */
    long long i0;                  /* REGULAR: first grid index */
    const long long* count_dev;    /* POINTS, optional (device): the actual number of points (<= count), read by the kernel --
                                      lets a compacted list be evaluated without a host synchronisation */
} gpb_segment;

typedef struct gpb_level {
    long long ld;                  /* leading dimension of the outputs (>= out_offset + count of every segment) */
    int n_segments;
    const gpb_segment* segments;   /* host [n_segments] */
    long long sp_offset;           /* output position of the surface-point tail (sp_all, n_sp_all points, evaluated by one
                                      of the segments), or -1; level 0 needs it */
    double* Z;                     /* [n_stacks][ld] */
    double* G;                     /* [n_stacks][3][ld] or NULL */
    double* block;                 /* [n_stacks][ld] activator output */
    double* final_block;           /* [ld] */
    double* faults_block;          /* [ld] */
    unsigned char* squeezed;       /* [n_stacks][ld] */
    unsigned char* mask;           /* [n_stacks][ld] or NULL */
    /* de-duplicated corners (optional, expand_map != NULL): the unique corners were evaluated at output positions
     * [expand_src, ...); gpb_model_combine first fills the corner segment [expand_dst, expand_dst + expand_count) of Z, G
     * and block with out[expand_dst + e] = out[expand_src + expand_map[e]], and combines m_combine points. */
    const int* expand_map;
    long long expand_src, expand_dst, expand_count;
    long long m_combine;           /* points the combination covers (0: up to the end of the last segment) */
} gpb_level;

typedef struct gpb_model gpb_model;
int  gpb_model_create(const gpb_model_desc* desc, gpb_model** out);
void gpb_model_destroy(gpb_model* m);
/* Assemble + solve stack i (weights, evaluation table, isovalues).  The earlier fault stacks must have been evaluated on
 * `level0` (their block rows and minima feed the fault-drift columns).  Synchronises the stream (reads the solver's
 * info); GPB_E_SINGULAR on a zero pivot.  path_host (may be NULL): 1 = symmetric path, 2 = pivoted LU. */
int  gpb_model_solve_stack(gpb_model* m, int i, const gpb_level* level0, int* path_host, void* workspace,
                           long long workspace_bytes, void* stream);
/* Scratch the largest system of the model needs (matrix + pivots).  Passing a device buffer of that size as `workspace`
 * avoids a stream-ordered allocation per solve (NULL / too small: the library allocates). */
long long gpb_model_workspace_bytes(const gpb_model* m);
/* Fused evaluation of stack i on every segment of the level: Z (+G), block, and for a fault stack the minimum of its
 * block in fault_min[i] (multi-GPU callers all-reduce it before the next dependent stack). */
int  gpb_model_eval_stack(gpb_model* m, int i, const gpb_level* lvl, void* stream);
int  gpb_model_combine(gpb_model* m, const gpb_level* lvl, void* stream);
/* solve (if `solve`) + evaluate every stack in order, then combine: the single-GPU path, one call per level. */
int  gpb_model_run_level(gpb_model* m, const gpb_level* lvl, int solve, void* workspace, long long workspace_bytes,
                         void* stream);
int  gpb_model_solver_path(const gpb_model* m, int i);

/* Corner de-duplication of an octree level: the 8 nvox corner slots of the voxel list (8 per voxel, the sign pattern of
 * gpb_voxel_corners) are grouped by lattice corner; _count returns the number of distinct corners (host; synchronises),
 * _emit writes their coordinates [3][ld_u] and map[8 nvox] (slot -> unique index; the representative of a group is its
 * first slot, so the result does not depend on scheduling).  Both calls share `scratch` (gpb_corner_scratch_bytes).
 * gpb_expand_rows: dst[r][e] = src[r][map[e]]. */
long long gpb_corner_scratch_bytes(long long nvox);
int gpb_corner_unique_count(const double* centers, long long ld_c, long long nvox, const gpb_regular_grid* lattice,
                            void* scratch, long long scratch_bytes, long long* n_unique_host, void* stream);
/* n_unique_host == NULL: no synchronisation, the count stays on the device.  A voxel list made of complete sibling
 * octets (every octree level below the root) has at most 27 distinct corners per 8 voxels, so buffers can be sized
 * without knowing the count.  count_dev_out (optional, device): receives count_base + number of unique corners. */
int gpb_corner_unique_emit(const double* centers, long long ld_c, long long nvox, double hx, double hy, double hz,
                           void* scratch, long long scratch_bytes, double* xyz_unique, long long ld_u, int* map,
                           long long count_base, long long* count_dev_out, void* stream);
int gpb_expand_rows(const double* src, long long ld_src, const int* map, int n_rows, long long count, double* dst,
                    long long ld_dst, void* stream);

/* dst[r][0..cols) = src[r][0..cols) for r < rows (device to device, on the copy engine). */
int gpb_copy_2d(double* dst, long long ld_dst, const double* src, long long ld_src, long long rows, long long cols, void* stream);

/* out[i] = rint(in[i]) (round half to even, as numpy.rint): the id blocks are rounded on the device before they are read
 * back (the reference rounds on the host: RawArraysSolution.lith_block = rint(final_block)). in may equal out. */
int gpb_rint(const double* in, long long n, double* out, void* stream);

/* ---- (4b) octree refinement  [engine stage "octrees_topology"] ----------------------------------- */
/* Corners of voxels (8 per voxel, sign pattern x:----++++ y:--++--++ z:-+-+-+-+), voxel-major:
 * corner = centre +- (hx, hy, hz); pass the half cell size. */
int gpb_voxel_corners(const double* centers, long long ld_c, long long nvox, double hx, double hy, double hz,
                      double* corners, long long ld_k, void* stream);
/* ids_corners: [8*nvox] combined lith+fault ids.  mark[v] = 1 if the 8 ids differ (or force_all). */
int gpb_mark_voxels(const double* lith_corners, const double* fault_corners, long long nvox, int force_all,
                    unsigned char* mark, void* stream);
/* Stable compaction of marked voxels into 8 children each (same sign pattern; child = centre +- (hx,hy,hz),
 * pass a quarter of the cell size).  children == NULL only counts.
 * Returns the number of children through n_children_host (host; this call synchronises the stream). */
int gpb_emit_children(const double* centers, long long ld_c, long long nvox, const unsigned char* mark,
                      double hx, double hy, double hz, double* children, long long ld_ch,
                      long long* n_children_host, void* stream);

/* The same in two steps without the second synchronisation: gpb_count_marked scans the marks into `offsets`
 * (gpb_scan_elems(nvox) long longs of device scratch) and returns the number of marked voxels (host; synchronises);
 * gpb_emit_marked then writes the 8 children of every marked voxel (asynchronous). */
long long gpb_scan_elems(long long nvox);
int gpb_count_marked(const unsigned char* mark, long long nvox, long long* offsets, long long* n_marked_host, void* stream);
int gpb_emit_marked(const double* centers, long long ld_c, long long nvox, const unsigned char* mark,
                    const long long* offsets, double hx, double hy, double hz, double* children, long long ld_ch,
                    void* stream);

/* Octree -> regular fill (RawArraysSolution.lith_block of an octree model): dst (2nx x 2ny x 2nz) = src (nx x ny x nz) with
 * every cell repeated twice per axis; then the level's own voxels overwrite their cells: dst[cell of centre v] = vals[v]
 * (rint'ed if round_ids), `lattice` = the level's voxel lattice (centre of cell (0,0,0) incl. shift, cell size, cells). */
int gpb_upsample2(const double* src, int nx, int ny, int nz, double* dst, void* stream);
int gpb_scatter_lattice(const double* centers, long long ld_c, long long nvox, const gpb_regular_grid* lattice,
                        const double* vals, int round_ids, double* dst, void* stream);

/* out[v] = 1 if any of in[8v .. 8v+7] is non-zero (voxel ownership from the squeezed mask at its corners). */
int gpb_any8(const unsigned char* in, long long nvox, unsigned char* out, void* stream);

/* ---- forward gravity  [engine stage "geophysics"; known answer test_gravity.py:89] ----------- */
/* out[c] = sum_k tz[k] * densities[id(c,k) - 1], id = rint(block[c * n_kernel + k]) clamped to 1..n_dens. */
int gpb_gravity(const double* block, const double* densities, int n_dens, const double* tz, int n_centers,
                long long n_kernel, double* out, void* stream);

/* ---- (4c) dual contouring  [engine stage "dual_contouring"] ---------------------------------------- */
/* Edge crossings of one isovalue: for voxel v and edge e (x x x x y y y y z z z z) valid[v*12+e] and the
 * crossing xyz_edge[3][12*nvox]; masked-out voxels (voxel_mask[v]==0) get no crossings. */
int gpb_dc_edges(const double* corners, long long ld_k, const double* Z_corners, long long nvox, double iso,
                 const unsigned char* voxel_mask, unsigned char* valid, double* xyz_edge, void* stream);
/* Per-voxel QEF (12 edge planes with raw gradient normals + 3 mass-point planes of strength `bias`).
 * grad_edge/xyz_edge: [3][12*nvox]; vertices: [3][nvox] (NaN for voxels without a crossing). */
int gpb_dc_vertices(const unsigned char* valid, const double* xyz_edge, const double* grad_edge, long long nvox,
                    double bias, double* vertices, void* stream);

/* One surface, entirely on the device and without a host synchronisation: crossings, stable compaction, gradient of the
 * stack's field at the compacted crossings (fused evaluation kernel), QEF vertices in voxel order, triangulation through a
 * hash table of the voxel lattice (two triangles per crossed edge shared by four surface voxels; (axis, voxel) order).
 * corners [3][ld_k] (8 per voxel), Z_corners [8 nvox], sq_corners [8 nvox] squeezed mask of the stack at the corners (NULL:
 * every voxel), centers [3][ld_c], iso_dev: the isovalue on the device, lattice: the level's voxel lattice (centre of cell
 * (0,0,0) with the shift, cell size, cells per axis).  Caller-allocated outputs sized for the worst case: valid [12 nvox],
 * xyz_c and grad_c [3][12 nvox] (compacted crossings / gradients; grad_c may be NULL), vertices [3][nvox] (compacted),
 * triangles [6 nvox][3], counts [3] = crossings, vertices, triangles (device).  scratch: gpb_dc_scratch_bytes(nvox). */
long long gpb_dc_scratch_bytes(long long nvox);
int gpb_dual_contour(const gpb_stack* st, const double* eval_table, const double* corners, long long ld_k,
                     const double* Z_corners, const unsigned char* sq_corners, const double* centers, long long ld_c,
                     long long nvox, const double* iso_dev, const gpb_regular_grid* lattice, double bias,
                     void* scratch, long long scratch_bytes, unsigned char* valid, double* xyz_c, double* grad_c,
                     double* vertices, int* triangles, long long* counts, void* stream);

/* ---- marching cubes on the dense grid (replaces the skimage.measure.marching_cubes call of
 * gempy/modules/mesh_extranction/marching_cubes.py:82-89) -------------------------------------------------
 * Z: scalar field on an nx*ny*nz lattice (x slowest, z fastest); mask: one byte per lattice point or NULL; a cube is
 * processed when the mask is set at its far corner (i+1, j+1, k+1).  "above" = Z > level.
 * Two calls: gpb_mc_count classifies (flags: m bytes; block_offsets: gpb_mc_scratch_elems(m) long longs), scans and
 * returns the totals through the two HOST pointers (synchronises the stream); gpb_mc_emit writes
 * vertices [V][3] = (index + t*axis) * (dx,dy,dz) + (ox,oy,oz) and triangles [T][3] (vertex ids, normals towards
 * lower values); vbase: m ints of scratch.  Vertex order: owner lattice point, then edge axis; triangle order:
 * cube, then case-table order. */
long long gpb_mc_scratch_elems(long long m);
int gpb_mc_count(const double* Z, const unsigned char* mask, int nx, int ny, int nz, double level,
                 unsigned char* flags, long long* block_offsets, long long* n_vertices_host,
                 long long* n_triangles_host, void* stream);
int gpb_mc_emit(const double* Z, const unsigned char* flags, const long long* block_offsets, int nx, int ny, int nz,
                double level, double ox, double oy, double oz, double dx, double dy, double dz, int* vbase,
                double* vertices, int* triangles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEMPY_B200_H */
