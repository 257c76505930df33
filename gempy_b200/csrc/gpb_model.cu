// Level executor: all stacks of a model on one evaluation domain, in native code.
//
// Engine stages replaced: interpolate_all_fields / interpolate_scalar_field of gempy_engine.compute_model (SURVEY.md 3.2,
// HOT LOOP 2; call site /root/reference/gempy/API/compute_API.py:68-73).  The reference walks the stacks in Python; so did
// round 1 of this backend, and 15 stacks x 8 octree levels cost ~800 ctypes calls and 60 ms of host time against 40 ms of
// device work.  Here one C call per level walks the stacks:
//
//   solve (level 0 only)   fault values of the earlier fault stacks at this stack's surface points -> covariance assembly
//                          -> symmetric solve (pivoted LU for tiny or indefinite systems) -> packed evaluation table
//                          -> scalar field at the stack's own surface points -> isovalues, their min / max
//   evaluate (every level) ONE fused launch per stack and segment: field (+ gradient) + fault drift read straight from the
//                          earlier stacks' block rows (minus their minima) + sigmoid activator + running minimum of the
//                          block (fault stacks)
//   combine                masks, top-down squeeze, lith / fault blocks
//
// All device buffers are the caller's (torch owns the memory); the handle keeps the description and small host tables.
#include "gpb_common.cuh"
#include <vector>
#include <cstring>

extern "C" int gpb_assemble_cov(const gpb_stack* st, double* A, int lda, double* b, void* stream);
extern "C" int gpb_assemble_cov_ex(const gpb_stack* st, double* A, int lda, double* b, int flags, void* stream);
extern "C" int gpb_sym_solve(int n, int nk, double* A, int lda, double* b, int nrhs, int ldb, int* info, void* stream);
extern "C" int gpb_lu_solve(int n, double* A, int lda, double* b, int nrhs, int ldb, int* ipiv, int* info, void* stream);
extern "C" int gpb_pack_eval_table(const gpb_stack* st, const double* w, double* src, void* stream);
extern "C" int gpb_expand_rows(const double* src, long long ld_src, const int* map, int n_rows, long long count, double* dst,
                               long long ld_dst, void* stream);
extern "C" int gpb_combine(const double* Z, const double* block, long long ld, long long m, int n_stacks,
                           const int* relations_host, const double* iso_min, const double* iso_max, double* final_block,
                           double* faults_block, unsigned char* squeezed_mask, unsigned char* mask, void* stream);

struct gpb_model {
    std::vector<gpb_model_stack> stacks;
    std::vector<std::vector<int>> faults_of;     // host copies of fault_stacks_host
    std::vector<int> rel;
    const double* sp_all = nullptr;
    long long n_sp_all = 0;
    double slope = 0.0;
    double* iso_min = nullptr;
    double* iso_max = nullptr;
    double* fault_min = nullptr;
    int solver = 0;
    std::vector<int> solved;                     // 0: not yet, 1: symmetric path, 2: LU
};

namespace {

constexpr int kSmallSystem = 160;                // one-CTA LU (gpb_lu.cu kSmallN)

// fault_rest[f][r] = block[g_f][sp_off + rest point r] - min[g_f], same for the reference point of increment r
__global__ void gather_fault_sp_kernel(int n_rest, int n_surf, int n_faults, const int* __restrict__ surf_offsets,
                                       const int* __restrict__ fault_ids, const double* __restrict__ block, long long ld,
                                       long long sp_off, const double* __restrict__ fault_min,
                                       double* __restrict__ fault_rest, double* __restrict__ fault_ref) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rest) return;
    int s = 0;
    while (s + 1 < n_surf && surf_offsets[s + 1] <= r) ++s;        // surface of increment r (n_surf <= 64)
    const long long i_rest = r + s + 1;                            // stack-local surface-point index: every surface starts with its reference point
    const long long i_ref = surf_offsets[s] + s;
    for (int f = 0; f < n_faults; ++f) {
        const int g = fault_ids[f];
        const double* row = block + (long long)g * ld + sp_off;
        const double mn = fault_min[g];
        fault_rest[(long long)f * n_rest + r] = row[i_rest] - mn;
        fault_ref[(long long)f * n_rest + r] = row[i_ref] - mn;
    }
}

// iso[s] = Z at the reference point of surface s; their min / max
__global__ void isovalues_kernel(int n_surf, const int* __restrict__ surf_offsets, const double* __restrict__ Zsp,
                                 double* __restrict__ iso, double* __restrict__ iso_min, double* __restrict__ iso_max) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double mn = __longlong_as_double(0x7ff0000000000000LL), mx = -mn;
    for (int s = 0; s < n_surf; ++s) {
        const double v = Zsp[surf_offsets[s] + s];
        iso[s] = v;
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    *iso_min = mn;
    *iso_max = mx;
}

int eval_segments(gpb_model* m, int i, const gpb_level* lvl, bool activate, cudaStream_t s) {
    const gpb_model_stack& ms = m->stacks[i];
    for (int k = 0; k < lvl->n_segments; ++k) {
        const gpb_segment& sg = lvl->segments[k];
        if (sg.count <= 0) continue;
        GpbEvalCall c;
        c.st = &ms.st;
        c.src = ms.eval_table;
        c.m = sg.count;
        if (sg.kind == GPB_SEG_REGULAR) {
            c.regular = 1;
            c.grid = sg.grid;
            c.i0 = sg.i0;
        } else {
            c.xyz = sg.xyz;
            c.ld_xyz = sg.ld_xyz;
            c.m_dev = sg.count_dev;
            c.octets = (sg.kind == GPB_SEG_OCTETS) ? 1 : 0;
        }
        const long long o = sg.out_offset;
        if (ms.st.n_faults > 0) {
            c.fault_vals = lvl->block + o;
            c.ld_fault = lvl->ld;
            c.fault_ids = ms.fault_stacks_dev;
            c.fault_min = m->fault_min;
        }
        c.Z = lvl->Z + (long long)i * lvl->ld + o;
        if (lvl->G != nullptr) {
            double* g = lvl->G + (long long)i * 3 * lvl->ld + o;
            c.gx = g;
            c.gy = g + lvl->ld;
            c.gz = g + 2 * lvl->ld;
        }
        if (activate) {
            c.block = lvl->block + (long long)i * lvl->ld + o;
            c.act_iso = ms.isovalues;
            c.act_ids = ms.unit_ids;
            c.act_n = ms.st.n_surf;
            c.act_slope = m->slope;
            c.block_min = (m->rel[i] == GPB_REL_FAULT) ? m->fault_min + i : nullptr;
        }
        int rc = gpb_eval_call(c, s);
        if (rc) return rc;
    }
    return GPB_OK;
}

}  // namespace

extern "C" int gpb_model_create(const gpb_model_desc* d, gpb_model** out) {
    GPB_REQUIRE(d && out && d->n_stacks >= 1 && d->n_stacks <= 64 && d->stacks, "bad model description (1 <= n_stacks <= 64)");
    GPB_REQUIRE(d->sp_all && d->iso_min && d->iso_max && d->fault_min, "null device buffers in the model description");
    gpb_model* m = new gpb_model();
    m->stacks.assign(d->stacks, d->stacks + d->n_stacks);
    m->faults_of.resize(d->n_stacks);
    m->rel.resize(d->n_stacks);
    m->solved.assign(d->n_stacks, 0);
    for (int i = 0; i < d->n_stacks; ++i) {
        gpb_model_stack& ms = m->stacks[i];
        m->rel[i] = ms.relation;
        bool ok = ms.st.n_faults >= 0 && (ms.st.n_faults == 0 || (ms.fault_stacks_host && ms.fault_stacks_dev && ms.st.fault_rest && ms.st.fault_ref)) &&
                  ms.weights && ms.eval_table && ms.isovalues && ms.unit_ids && ms.st.n_surf >= 1 && ms.st.n_surf <= 64;
        for (int f = 0; ok && f < ms.st.n_faults; ++f) {
            const int g = ms.fault_stacks_host[f];
            ok = g >= 0 && g < i;                                   // only earlier stacks can drift this one
            m->faults_of[i].push_back(g);
        }
        if (!ok) {
            delete m;
            return gpb_set_error(GPB_E_INVALID, "stack %d: incomplete description (buffers, 1 <= n_surf <= 64, faults must be earlier stacks)", i);
        }
        ms.fault_stacks_host = nullptr;                             // the caller's host array need not outlive this call
    }
    m->sp_all = d->sp_all;
    m->n_sp_all = d->n_sp_all;
    m->slope = d->sigmoid_slope;
    m->iso_min = d->iso_min;
    m->iso_max = d->iso_max;
    m->fault_min = d->fault_min;
    m->solver = d->solver;
    *out = m;
    return GPB_OK;
}

extern "C" void gpb_model_destroy(gpb_model* m) { delete m; }

extern "C" long long gpb_model_workspace_bytes(const gpb_model* m) {
    long long best = 0;
    if (!m) return 0;
    for (const gpb_model_stack& ms : m->stacks) {
        const long long n = 3LL * ms.st.n_ori + ms.st.n_rest + ms.st.n_drift + ms.st.n_faults;
        const long long lda = (n + 2) & ~1LL;
        const long long b = 8 * lda * n + 4 * (n + 4);
        if (b > best) best = b;
    }
    return best;
}

extern "C" int gpb_model_solve_stack(gpb_model* m, int i, const gpb_level* level0, int* path_host, void* workspace,
                                     long long workspace_bytes, void* stream) {
    GPB_REQUIRE(m && i >= 0 && i < (int)m->stacks.size() && level0, "bad arguments");
    GPB_REQUIRE(level0->sp_offset >= 0 && level0->Z && level0->block, "level 0 must carry the surface-point tail");
    cudaStream_t s = (cudaStream_t)stream;
    gpb_model_stack& ms = m->stacks[i];
    const gpb_stack& st = ms.st;
    const long long sp_off = level0->sp_offset + ms.sp_begin;
    if (st.n_faults > 0 && st.n_rest > 0) {
        gather_fault_sp_kernel<<<(st.n_rest + 127) / 128, 128, 0, s>>>(st.n_rest, st.n_surf, st.n_faults, st.surf_offsets, ms.fault_stacks_dev,
                                                                       level0->block, level0->ld, sp_off, m->fault_min,
                                                                       const_cast<double*>(st.fault_rest), const_cast<double*>(st.fault_ref));
        GPB_LAUNCH_CHECK();
    }
    const int n = 3 * st.n_ori + st.n_rest + st.n_drift + st.n_faults;
    const int nk = 3 * st.n_ori + st.n_rest;
    GPB_REQUIRE(n >= 1, "empty system");
    const bool try_sym = m->solver == 0 && n > kSmallSystem && nk >= 1 && (n - nk) <= 64;
    const int lda = (n + 2) & ~1;                                   // even, >= n + 1: room for the right-hand side row of the symmetric path
    char* ws = nullptr;
    const size_t a_bytes = sizeof(double) * (size_t)lda * n;
    // the caller's workspace when it is large enough (a 9.8 GB stream-ordered allocation costs more than the solve it is
    // for: the pool gives the memory back at every synchronisation), else a stream-ordered allocation
    const bool own_ws = !(workspace != nullptr && (size_t)workspace_bytes >= a_bytes + sizeof(int) * (size_t)(n + 4));
    if (own_ws) GPB_CHECK_CUDA(gpb_malloc_async((void**)&ws, a_bytes + sizeof(int) * (size_t)(n + 4), s));
    else ws = (char*)workspace;
    double* A = reinterpret_cast<double*>(ws);
    int* ipiv = reinterpret_cast<int*>(ws + a_bytes);
    int* info = ipiv + n;
    int path = 0, rc = GPB_OK, info_h = 0;
    if (try_sym) {
        rc = gpb_assemble_cov_ex(&st, A, lda, ms.weights, GPB_COV_LOWER_ONLY, s);
        if (!rc) rc = gpb_sym_solve(n, nk, A, lda, ms.weights, 1, n, info, s);
        if (!rc && cudaMemcpyAsync(&info_h, info, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = gpb_set_error(GPB_E_CUDA, "info copy failed");
        if (!rc && cudaStreamSynchronize(s) != cudaSuccess) rc = gpb_set_error(GPB_E_CUDA, "stream synchronisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (!rc && info_h == 0) path = 1;
    }
    if (!rc && path == 0) {                                          // tiny, forced or not positive definite: pivoted LU
        rc = gpb_assemble_cov(&st, A, lda, ms.weights, s);
        if (!rc) rc = gpb_lu_solve(n, A, lda, ms.weights, 1, n, ipiv, info, s);
        if (!rc && cudaMemcpyAsync(&info_h, info, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = gpb_set_error(GPB_E_CUDA, "info copy failed");
        if (!rc && cudaStreamSynchronize(s) != cudaSuccess) rc = gpb_set_error(GPB_E_CUDA, "stream synchronisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (!rc && info_h != 0)
            rc = gpb_set_error(GPB_E_SINGULAR, "stack %d: zero pivot at column %d -- the co-kriging system is singular (duplicate data with zero nugget, or an all-zero drift column)", i, info_h);
        path = 2;
    }
    if (own_ws) cudaFreeAsync(ws, s);
    if (rc) return rc;
    m->solved[i] = path;
    if (path_host) *path_host = path;
    if ((rc = gpb_pack_eval_table(&st, ms.weights, ms.eval_table, s))) return rc;
    // scalar field at the stack's own surface points (written where the level-0 evaluation will write the same values)
    GpbEvalCall c;
    c.st = &st;
    c.src = ms.eval_table;
    c.xyz = m->sp_all + ms.sp_begin;
    c.ld_xyz = m->n_sp_all;
    c.m = ms.n_sp;
    if (st.n_faults > 0) {
        c.fault_vals = level0->block + sp_off;
        c.ld_fault = level0->ld;
        c.fault_ids = ms.fault_stacks_dev;
        c.fault_min = m->fault_min;
    }
    double* Zsp = level0->Z + (long long)i * level0->ld + sp_off;
    c.Z = Zsp;
    if ((rc = gpb_eval_call(c, s))) return rc;
    isovalues_kernel<<<1, 32, 0, s>>>(st.n_surf, st.surf_offsets, Zsp, ms.isovalues, m->iso_min + i, m->iso_max + i);
    GPB_LAUNCH_CHECK();
    return GPB_OK;
}

extern "C" int gpb_model_eval_stack(gpb_model* m, int i, const gpb_level* lvl, void* stream) {
    GPB_REQUIRE(m && i >= 0 && i < (int)m->stacks.size() && lvl && lvl->Z && lvl->block && lvl->segments, "bad arguments");
    GPB_REQUIRE(m->solved[i] != 0, "stack not solved yet (gpb_model_solve_stack)");
    cudaStream_t s = (cudaStream_t)stream;
    if (m->rel[i] == GPB_REL_FAULT) GPB_CHECK_CUDA(cudaMemsetAsync(m->fault_min + i, 0x7f, sizeof(double), s));      // 1.4e306
    return eval_segments(m, i, lvl, true, s);
}

extern "C" int gpb_model_combine(gpb_model* m, const gpb_level* lvl, void* stream) {
    GPB_REQUIRE(m && lvl && lvl->final_block && lvl->faults_block && lvl->squeezed, "bad arguments");
    const int n_st = (int)m->stacks.size();
    if (lvl->expand_map != nullptr && lvl->expand_count > 0) {       // corner segment <- unique-corner results
        int rc = gpb_expand_rows(lvl->Z + lvl->expand_src, lvl->ld, lvl->expand_map, n_st, lvl->expand_count, lvl->Z + lvl->expand_dst, lvl->ld, stream);
        if (!rc) rc = gpb_expand_rows(lvl->block + lvl->expand_src, lvl->ld, lvl->expand_map, n_st, lvl->expand_count, lvl->block + lvl->expand_dst, lvl->ld, stream);
        if (!rc && lvl->G != nullptr)
            rc = gpb_expand_rows(lvl->G + lvl->expand_src, lvl->ld, lvl->expand_map, 3 * n_st, lvl->expand_count, lvl->G + lvl->expand_dst, lvl->ld, stream);
        if (rc) return rc;
    }
    long long mtot = lvl->m_combine;
    if (mtot <= 0)
        for (int k = 0; k < lvl->n_segments; ++k) {
            const long long e = lvl->segments[k].out_offset + lvl->segments[k].count;
            if (e > mtot) mtot = e;
        }
    return gpb_combine(lvl->Z, lvl->block, lvl->ld, mtot, n_st, m->rel.data(), m->iso_min, m->iso_max,
                       lvl->final_block, lvl->faults_block, lvl->squeezed, lvl->mask, stream);
}

extern "C" int gpb_model_run_level(gpb_model* m, const gpb_level* lvl, int solve, void* workspace, long long workspace_bytes,
                                   void* stream) {
    GPB_REQUIRE(m && lvl, "bad arguments");
    for (int i = 0; i < (int)m->stacks.size(); ++i) {
        int rc;
        if (solve && (rc = gpb_model_solve_stack(m, i, lvl, nullptr, workspace, workspace_bytes, stream))) return rc;
        if ((rc = gpb_model_eval_stack(m, i, lvl, stream))) return rc;
    }
    return gpb_model_combine(m, lvl, stream);
}

extern "C" int gpb_model_solver_path(const gpb_model* m, int i) {
    if (!m || i < 0 || i >= (int)m->stacks.size()) return -1;
    return m->solved[i];
}
