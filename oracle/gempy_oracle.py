"""CPU oracle for the B200 backend -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package ``gempy_b200`` never does.

What it is
----------
A plain numpy float64 restatement of the algorithm GemPy hands to ``gempy_engine.compute_model``
(call site /root/reference/gempy/API/compute_API.py:68-73).  The arithmetic lives in the third-party
package ``gempy_engine>=2026.0.3`` (/root/reference/requirements/requirements.txt:2) which is NOT vendored
in the reference tree, not installed and not installable here (no network) -- so nothing below could be
compared with a run of the real engine.  The restatement follows

  * the published method: de la Varga, Schaaf & Wellmann (2019), GMD 12, 1-32 (universal co-kriging of a
    potential field; cubic covariance; Lajaunie et al. 1997 increments formulation);
  * the conventions the reference tree itself pins: kernel_options defaults (range 1.7, c_o 10, i_res 4,
    gi_res 2, uni_degree 1, cubic) and evaluation/octree defaults in
    test/test_modules/test_serialize_model.*.verify/*.approved.txt; default nuggets
    (gempy/core/data/surface_points.py:11 = 2e-5, orientations.py:11 = 0.01); reference point = first point
    of each surface; grid ordering gempy/core/data/grid_modules/regular_grid.py:58-71; unit ids
    structural_frame.py:367-370; stack relations structural_frame.py:325-330; fault matrix 243-273;
  * the upstream engine's structure as publicly documented (block scaling i_res/gi_res inherited from
    GemPy v2, ``h_u h_v / (r^2 + 1e-5)`` regulariser in the gradient-gradient block, sigmoid activator,
    ERODE/FAULT/BASEMENT masks, octree by corner ids + scalar statistics, dual contouring with a
    mass-point bias of strength 1).

PARITY STATUS: PINNED against the reference's own golden vectors (tests/test_oracle.py):
  * all three ACTIVE approved scalar-field vectors of test/test_model_types/test_example_models_I.py
    (Anticline :41-52, Fault :55-66, Combination :69-88 -- 51 values each, sampled
    ``octrees_output[-1].outputs[0].exported_fields.scalar_field[::len//50]``) are reproduced 51/51 to
    < 1e-7 (the reference's own tolerance is 1e-5).  They pin, jointly: the cubic kernel and its block scaling,
    the 1e-5 regulariser, the 1e-10 distance epsilon, nuggets inside the c_o factor, the +1e-6 grid shift,
    the reference-point convention, the octree leaf order and the corner-id refinement rule, and -- through
    the leaf lists, which depend on the lithology/fault ids of every corner at every level -- the fault
    drift, activator, ERODE/FAULT/BASEMENT masks and the stack combination.
  * the ENGINE OUTPUTS stored in the header of examples/data/gempy_models/Greenstone.gempy (the model of
    gempy/API/examples_generator.py:489-508: 3 series, 70 surface points, 41 orientations): the scalar field at the
    four interfaces, 16 digits each, reproduced to 2e-12 / 3e-13 absolute.  This pins at full double precision the
    assembly with 26 orientations in one stack, the nuggets, the drift, the solve and the evaluation at the surface
    points (which are NOT shifted by the 1e-6 the grid points get), and `Transform.from_input_points`.
  * custom-grid lith ids [3,3,3,3,1,1,1,1] (test/test_modules/test_grids/test_custom_grid.py:44-47).
  * HORIZONTAL_STRAT's approved vector is disabled in the reference (test_example_models_I.py:34 ``if False``)
    and stale; the plane solution Z = gi_res * z' is asserted instead.
NOT pinned (no fixture in the reference): non-uniform nuggets, universal degree 2, exponential / Matern
kernels, ONLAP, the exported-gradient scaling, dual-contouring vertices.  See DESIGN.md "Oracle pinning".

Conventions (all in the transformed coordinate system the bridge hands over)
-----------------------------------------------------------------------------
System rows/cols: [G_x(1..n_o), G_y, G_z | rest_i - ref_i (n_rest) | universal drift (n_u) | faults (n_f)].
  C_GG[(o,a),(p,b)] = c_o * ( h_a h_b/(r^2+1e-5) * (C'/r - C'') - delta_ab C'/r ),  h = x_o - x_p
  C_GI[(o,a), i]    = c_o * gi_res * ( (x_o-rest_i)_a C'/r|rest - (x_o-ref_i)_a C'/r|ref )
  C_II[i, j]        = c_o * i_res * ( C(rest_i,rest_j) - C(rest_i,ref_j) - C(ref_i,rest_j) + C(ref_i,ref_j) )
  U_G[(o,a), k]     = d f_k / d x_a (x_o),        U_I[i,k] = gi_res * (f_k(rest_i) - f_k(ref_i))
  F_G = 0,                                       F_I[i,f] = F_f(rest_i) - F_f(ref_i)
  diag += c_o * [nugget_grad x3 | (nugget_rest + nugget_ref)/2 | 0 | 0];    b = [G_x, G_y, G_z, 0...]
  every distance is r = sqrt(|h|^2 + 1e-10); regular-grid and octree points are evaluated at centre + 1e-6
Evaluation (same block scaling, the grid point playing the role of an interface point):
  Z(x)   = c_o*( gi_res * sum_(o,a) w_oa (x_o - x)_a C'/r  +  i_res * sum_i w_i (C(x,rest_i) - C(x,ref_i)) )
           + gi_res * sum_k mu_k f_k(x) + sum_f w_f F_f(x)
  dZ/dx_a (engine "gradient kernel" convention: the grid point plays the role of an orientation, so the
  exported gradient is (1/gi_res) * the analytic derivative of Z, with the same 1e-5 regulariser):
  G_a(x) = c_o*( sum_(o,b) w_ob ( h_a h_b/(r^2+1e-5) (C'/r - C'') - delta_ab C'/r ),  h = x - x_o
                 + gi_res * sum_i w_i ( (x-rest_i)_a C'/r|rest - (x-ref_i)_a C'/r|ref ) )
           + sum_k mu_k d f_k/d x_a (x)
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np

REG_EPS = 1e-5        # regulariser of the gradient-gradient term              (pinned: Anticline golden)
DIST_EPS = 1e-10      # r = sqrt(|h|^2 + DIST_EPS)                             (pinned: Anticline golden)
GRID_SHIFT = 1e-6     # regular-grid / octree points sit at centre + 1e-6      (pinned: Combination + Anticline)
# The approved vectors are reproduced 51/51 with the corner-id test alone (levels below octree_min_level
# fully refined); adding a scalar-statistics test changes the leaf lists and breaks them, so it is off.
USE_STATS_REFINEMENT = False
ERODE, ONLAP, FAULT, BASEMENT = 1, 2, 3, 4


def _rel_code(rel) -> int:
    if rel is False or rel is None:
        return BASEMENT
    return int(getattr(rel, "value", rel))


# ----------------------------------------------------------------------------------------------
# covariance functions: C(r), C'(r)/r, C''(r), all for c_o = 1   (SURVEY.md §8c "Kernel definitions")
# ----------------------------------------------------------------------------------------------
def _kernel_name(k) -> str:
    return getattr(k, "name", k)


def kernel_terms(r: np.ndarray, a: float, kind="cubic"):
    """Return (C, C'/r, C'') at distance r for range a.
    Follows: engine kernel_functions (absent from the tree); "kernel_function": "cubic" is the only kernel the
    reference's fixtures select (test/test_modules/test_serialize_model.*.verify/*.approved.txt, kernel_options);
    cubic pinned by the approved vectors and the Greenstone isovalues, exponential / Matern unpinned."""
    kind = _kernel_name(kind)
    if kind == "cubic":
        t = r / a
        t2 = t * t
        C = 1 - 7 * t2 + 35 / 4 * t2 * t - 7 / 2 * t2 * t2 * t + 3 / 4 * t2 * t2 * t2 * t
        Cp_r = (-14 + 105 / 4 * t - 35 / 2 * t2 * t + 21 / 4 * t2 * t2 * t) / a ** 2
        Cpp = 7 * (9 * t2 * t2 * t - 20 * t2 * t + 15 * t - 4) / (2 * a ** 2)
        return C, Cp_r, Cpp
    if kind == "exponential":      # upstream's "exponential" is the Gaussian-type exp(-r^2 / (2 a^2))
        e = np.exp(-(r * r) / (2 * a * a))
        return e, -e / a ** 2, e * (r * r / a ** 4 - 1 / a ** 2)
    if kind == "matern_5_2":
        s = np.sqrt(5.0) * r / a
        e = np.exp(-s)
        C = (1 + s + s * s / 3) * e
        Cp_r = -(5.0 / (3 * a * a)) * (1 + s) * e
        Cpp = -(5.0 / (3 * a * a)) * (1 + s - s * s) * e
        return C, Cp_r, Cpp
    raise ValueError(f"unknown kernel {kind}")


# ----------------------------------------------------------------------------------------------
# universal drift basis
# ----------------------------------------------------------------------------------------------
def n_drift_terms(degree: int) -> int:
    return {0: 0, 1: 3, 2: 9}[int(degree)]


def drift_basis(x: np.ndarray, degree: int) -> np.ndarray:
    """f_k(x): (m, n_u). Order: x, y, z, x^2, y^2, z^2, xy, xz, yz."""
    if degree == 0:
        return np.zeros((x.shape[0], 0))
    cols = [x[:, 0], x[:, 1], x[:, 2]]
    if degree == 2:
        cols += [x[:, 0] ** 2, x[:, 1] ** 2, x[:, 2] ** 2, x[:, 0] * x[:, 1], x[:, 0] * x[:, 2], x[:, 1] * x[:, 2]]
    return np.stack(cols, axis=1)


def drift_basis_grad(x: np.ndarray, degree: int) -> np.ndarray:
    """d f_k / d x_a: (3, m, n_u)."""
    m = x.shape[0]
    nu = n_drift_terms(degree)
    g = np.zeros((3, m, nu))
    if degree >= 1:
        for a in range(3):
            g[a, :, a] = 1.0
    if degree == 2:
        for a in range(3):
            g[a, :, 3 + a] = 2 * x[:, a]
        g[0, :, 6] = x[:, 1]; g[1, :, 6] = x[:, 0]
        g[0, :, 7] = x[:, 2]; g[2, :, 7] = x[:, 0]
        g[1, :, 8] = x[:, 2]; g[2, :, 8] = x[:, 1]
    return g


# ----------------------------------------------------------------------------------------------
# one stack (= one scalar field)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class StackData:
    ref: np.ndarray            # (n_rest, 3) reference point of each rest point's surface
    rest: np.ndarray           # (n_rest, 3)
    nugget_sp: np.ndarray      # (n_rest,)  nugget_rest + nugget_ref
    ori_pos: np.ndarray        # (n_o, 3)
    ori_grad: np.ndarray       # (n_o, 3)
    nugget_ori: np.ndarray     # (n_o,)
    ref_idx: np.ndarray        # (n_surf,) index of each surface's reference point inside the stack's sp table
    sp_all: np.ndarray         # (n_sp, 3)
    n_per_surface: np.ndarray  # (n_surf,)
    fault_ref: np.ndarray      # (n_rest, n_f)
    fault_rest: np.ndarray     # (n_rest, n_f)

    @property
    def n_o(self): return self.ori_pos.shape[0]

    @property
    def n_rest(self): return self.rest.shape[0]

    @property
    def n_f(self): return self.fault_ref.shape[1]


def prepare_stack(sp, sp_nugget, n_per_surface, ori_pos, ori_grad, ori_nugget, fault_on_sp=None) -> StackData:
    """ref/rest split: the first point of each surface is its reference point.
    Inputs as built by gempy/modules/data_manipulation/_engine_factory.py:26-37 (SurfacePoints / Orientations) and
    gempy/core/data/structural_frame.py:333-350 (points per element / group); point order = group order, then element
    order, then table order (structural_frame.py:377-381)."""
    sp = np.asarray(sp, float).reshape(-1, 3)
    n_per_surface = np.asarray(n_per_surface, int)
    starts = np.concatenate([[0], np.cumsum(n_per_surface)[:-1]]).astype(int)
    is_ref = np.zeros(sp.shape[0], bool)
    is_ref[starts] = True
    reps = n_per_surface - 1
    ref = np.repeat(sp[starts], reps, axis=0)
    rest = sp[~is_ref]
    sp_nugget = np.broadcast_to(np.asarray(sp_nugget, float), (sp.shape[0],))
    # Row nugget of an increment rest_i - ref_i.  The goldens pin c_o * 2e-5 per row when every point
    # carries the default 2e-5; for non-uniform nuggets the mean of the pair is this oracle's choice.
    nug = 0.5 * (sp_nugget[~is_ref] + np.repeat(sp_nugget[starts], reps))
    if fault_on_sp is None:
        fault_on_sp = np.zeros((0, sp.shape[0]))
    fault_on_sp = np.asarray(fault_on_sp, float).reshape(-1, sp.shape[0])
    f_ref = np.repeat(fault_on_sp[:, starts], reps, axis=1).T
    f_rest = fault_on_sp[:, ~is_ref].T
    ori_pos = np.asarray(ori_pos, float).reshape(-1, 3)
    return StackData(ref, rest, nug, ori_pos, np.asarray(ori_grad, float).reshape(-1, 3),
                     np.broadcast_to(np.asarray(ori_nugget, float), (ori_pos.shape[0],)).copy(),
                     starts, sp, n_per_surface, f_ref.reshape(rest.shape[0], -1), f_rest.reshape(rest.shape[0], -1))


def _dist(a, b):
    d = a[:, None, :] - b[None, :, :]
    return d, np.sqrt((d * d).sum(-1) + DIST_EPS)


def system_size(st: StackData, ko) -> int:
    return 3 * st.n_o + st.n_rest + n_drift_terms(ko.uni_degree) + st.n_f


def assemble_covariance(st: StackData, ko) -> np.ndarray:
    """The saddle-point matrix of one stack (engine stage "kernel_constructor"; call site
    gempy/API/compute_API.py:68-73).  Every constant here is pinned by test/test_model_types/*.approved.txt and by the
    isovalues in examples/data/gempy_models/Greenstone.gempy (see the module header)."""
    a, c_o, gi, ires, kind = ko.range, ko.c_o, ko.gi_res, ko.i_res, ko.kernel_function
    n_o, n_r = st.n_o, st.n_rest
    nu = n_drift_terms(ko.uni_degree)
    n = 3 * n_o + n_r + nu + st.n_f
    A = np.zeros((n, n))
    # --- C_GG
    if n_o:
        h, r = _dist(st.ori_pos, st.ori_pos)
        _, kp, ka = kernel_terms(r, a, kind)
        T = (kp - ka) / (r * r + REG_EPS)
        for ia in range(3):
            for ib in range(3):
                blk = h[:, :, ia] * h[:, :, ib] * T
                if ia == ib:
                    blk = blk - kp
                A[ia * n_o:(ia + 1) * n_o, ib * n_o:(ib + 1) * n_o] = c_o * blk
    # --- C_GI
    if n_o and n_r:
        h_rest, r_rest = _dist(st.ori_pos, st.rest)
        h_ref, r_ref = _dist(st.ori_pos, st.ref)
        _, kp_rest, _ = kernel_terms(r_rest, a, kind)
        _, kp_ref, _ = kernel_terms(r_ref, a, kind)
        for ia in range(3):
            blk = c_o * gi * (h_rest[:, :, ia] * kp_rest - h_ref[:, :, ia] * kp_ref)
            A[ia * n_o:(ia + 1) * n_o, 3 * n_o:3 * n_o + n_r] = blk
            A[3 * n_o:3 * n_o + n_r, ia * n_o:(ia + 1) * n_o] = blk.T
    # --- C_II
    if n_r:
        k = lambda p, q: kernel_terms(_dist(p, q)[1], a, kind)[0]
        blk = k(st.rest, st.rest) - k(st.rest, st.ref) - k(st.ref, st.rest) + k(st.ref, st.ref)
        A[3 * n_o:3 * n_o + n_r, 3 * n_o:3 * n_o + n_r] = c_o * ires * blk
    # --- universal drift
    o_u = 3 * n_o + n_r
    if nu:
        ug = drift_basis_grad(st.ori_pos, ko.uni_degree)            # (3, n_o, nu)
        for ia in range(3):
            A[ia * n_o:(ia + 1) * n_o, o_u:o_u + nu] = ug[ia]
            A[o_u:o_u + nu, ia * n_o:(ia + 1) * n_o] = ug[ia].T
        ui = gi * (drift_basis(st.rest, ko.uni_degree) - drift_basis(st.ref, ko.uni_degree))
        A[3 * n_o:3 * n_o + n_r, o_u:o_u + nu] = ui
        A[o_u:o_u + nu, 3 * n_o:3 * n_o + n_r] = ui.T
    # --- fault drift
    o_f = o_u + nu
    if st.n_f:
        fi = st.fault_rest - st.fault_ref
        A[3 * n_o:3 * n_o + n_r, o_f:] = fi
        A[o_f:, 3 * n_o:3 * n_o + n_r] = fi.T
    # --- nugget
    d = np.concatenate([np.tile(st.nugget_ori, 3), st.nugget_sp, np.zeros(nu + st.n_f)])
    A[np.diag_indices(n)] += c_o * d          # nuggets sit inside the c_o factor (pinned: Anticline golden)
    return A


def rhs(st: StackData, ko) -> np.ndarray:
    b = np.zeros(system_size(st, ko))
    b[:3 * st.n_o] = st.ori_grad.T.ravel()        # [G_x.., G_y.., G_z..]
    return b


def solve(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    """kernel_solver = 1 (direct dense solve, serialization golden kernel_options.kernel_solver): LAPACK gesv."""
    return np.linalg.solve(A, b)


def evaluate(st: StackData, ko, w: np.ndarray, xyz: np.ndarray, fault_at_xyz: Optional[np.ndarray] = None,
             gradient: bool = False, chunk_elems: int = 4_000_000):
    """Z (m,) and, if requested, the engine-convention gradient (m,3) at xyz."""
    a, c_o, gi, ires, kind = ko.range, ko.c_o, ko.gi_res, ko.i_res, ko.kernel_function
    n_o, n_r = st.n_o, st.n_rest
    nu = n_drift_terms(ko.uni_degree)
    w_g = w[:3 * n_o].reshape(3, n_o)              # [a, o]
    w_i = w[3 * n_o:3 * n_o + n_r]
    mu = w[3 * n_o + n_r:3 * n_o + n_r + nu]
    w_f = w[3 * n_o + n_r + nu:]
    xyz = np.asarray(xyz, float).reshape(-1, 3)
    m = xyz.shape[0]
    Z = np.zeros(m)
    G = np.zeros((m, 3)) if gradient else None
    step = max(1, int(chunk_elems // max(1, (3 * n_o + 2 * n_r))))
    for s in range(0, m, step):
        x = xyz[s:s + step]
        z = np.zeros(x.shape[0])
        g = np.zeros((x.shape[0], 3)) if gradient else None
        if n_o:
            h, r = _dist(x, st.ori_pos)            # h = x - x_o
            _, kp, ka = kernel_terms(r, a, kind)
            hw = np.einsum("moa,ao->mo", h, w_g)   # h . w_o
            z += c_o * gi * (-(hw * kp)).sum(1)    # (x_o - x)_a = -h_a
            if gradient:
                T = (kp - ka) / (r * r + REG_EPS)
                g += c_o * (np.einsum("moa,mo->ma", h, hw * T) - np.einsum("mo,ao->ma", kp, w_g))
        if n_r:
            h1, r1 = _dist(x, st.rest)
            h0, r0 = _dist(x, st.ref)
            C1, kp1, _ = kernel_terms(r1, a, kind)
            C0, kp0, _ = kernel_terms(r0, a, kind)
            z += c_o * ires * ((C1 - C0) @ w_i)
            if gradient:
                g += c_o * gi * (np.einsum("mia,mi->ma", h1, kp1 * w_i) - np.einsum("mia,mi->ma", h0, kp0 * w_i))
        if nu:
            z += gi * (drift_basis(x, ko.uni_degree) @ mu)
            if gradient:
                g += np.einsum("amk,k->ma", drift_basis_grad(x, ko.uni_degree), mu)
        if w_f.size:
            z += fault_at_xyz[:, s:s + step].T @ w_f
        Z[s:s + step] = z
        if gradient:
            G[s:s + step] = g
    return (Z, G) if gradient else Z


# ----------------------------------------------------------------------------------------------
# activator: sum of steep sigmoids mapping Z to unit ids
# ----------------------------------------------------------------------------------------------
def _sig(x):
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


def activate(Z: np.ndarray, isovalues: np.ndarray, ids: np.ndarray, slope: float) -> np.ndarray:
    """Engine stage "activator" (SURVEY.md 8a2 row 4a); sigmoid_slope = 5e6 from the serialization golden; unit ids from
    gempy/core/data/structural_frame.py:367-370; known answer test/test_modules/test_grids/test_custom_grid.py:44-47.
    block(Z) = sum_k ids[k] * (sigma(l (Z - lower_k)) - sigma(l (Z - upper_k))).
    Interval k lies between isovalues[k-1] (upper) and isovalues[k] (lower); the first has no upper
    bound, the last no lower bound.  ids has len(isovalues)+1 entries."""
    iso = np.asarray(isovalues, float)
    ids = np.asarray(ids, float)
    n = iso.shape[0]
    out = np.zeros_like(Z)
    for k in range(n + 1):
        lower = _sig(slope * (Z - iso[k])) if k < n else 1.0
        upper = _sig(slope * (Z - iso[k - 1])) if k > 0 else 0.0
        out += ids[k] * (lower - upper)
    return out


# ----------------------------------------------------------------------------------------------
# all stacks of a model on one point set
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class StackResult:
    weights: np.ndarray
    Z: np.ndarray                 # (n_xyz,) on grid ++ all surface points
    G: Optional[np.ndarray]
    isovalues: np.ndarray
    values_block: np.ndarray
    relation: int
    mask: np.ndarray = None
    squeezed_mask: np.ndarray = None
    cond: Optional[float] = None


@dataclasses.dataclass
class FieldsResult:
    stacks: List[StackResult]
    final_block: np.ndarray
    faults_block: np.ndarray
    grid_size: int

    @property
    def lith_ids(self):
        return np.rint(self.final_block[:self.grid_size])

    @property
    def litho_faults_ids(self):
        lith = np.rint(self.final_block[:self.grid_size])
        f = np.rint(self.faults_block[:self.grid_size])
        return lith + f * max(len(np.unique(lith)), 1)


def _stack_slices(descriptor):
    ss, ts = descriptor.stack_structure, descriptor.tensors_structure
    sp0 = np.concatenate([[0], np.cumsum(ss.number_of_points_per_stack)]).astype(int)
    or0 = np.concatenate([[0], np.cumsum(ss.number_of_orientations_per_stack)]).astype(int)
    su0 = np.concatenate([[0], np.cumsum(ss.number_of_surfaces_per_stack)]).astype(int)
    return sp0, or0, su0


def interpolate_all_fields(interp_input, options, descriptor, xyz_grid: np.ndarray,
                           weights_cache: Optional[list] = None, gradient: Optional[bool] = None) -> FieldsResult:
    """Per stack: subset -> (assemble, solve | cached weights) -> evaluate on grid++surface points ->
    activator -> mask; then combine top-down."""
    ko = options.kernel_options
    if gradient is None:
        gradient = options.evaluation_options.compute_scalar_gradient
    ss, ts = descriptor.stack_structure, descriptor.tensors_structure
    sp0, or0, su0 = _stack_slices(descriptor)
    sp_all = interp_input.surface_points.sp_coords
    xyz = np.vstack([np.asarray(xyz_grid, float).reshape(-1, 3), sp_all])
    gsz = xyz.shape[0] - sp_all.shape[0]
    n_st = ss.n_stacks
    unit_values = np.asarray(interp_input.unit_values, float)
    rel = [_rel_code(r) for r in ss.masking_descriptor]
    fr = ss.faults_relations if ss.faults_relations is not None else np.zeros((n_st, n_st), bool)
    values_everywhere = np.zeros((n_st, xyz.shape[0]))
    results: List[StackResult] = []
    for i in range(n_st):
        sl_sp, sl_or = slice(sp0[i], sp0[i + 1]), slice(or0[i], or0[i + 1])
        active = np.nonzero(np.asarray(fr)[:, i])[0]
        f_every = values_everywhere[active]                                     # (n_f, n_xyz)
        f_on_sp = f_every[:, gsz:][:, sl_sp]
        st = prepare_stack(sp_all[sl_sp], interp_input.surface_points.nugget_effect_scalar[sl_sp],
                           ts.number_of_points_per_surface[su0[i]:su0[i + 1]],
                           interp_input.orientations.dip_positions[sl_or],
                           interp_input.orientations.dip_gradients[sl_or],
                           interp_input.orientations.nugget_effect_grad[sl_or], f_on_sp)
        cond = None
        if weights_cache is not None and weights_cache[i] is not None:
            w = weights_cache[i]
        else:
            A = assemble_covariance(st, ko)
            if getattr(ko, "compute_condition_number", False):
                cond = float(np.linalg.cond(A))
            w = solve(A, rhs(st, ko))
            if weights_cache is not None:
                weights_cache[i] = w
        out = evaluate(st, ko, w, xyz, f_every, gradient=gradient)
        Z, G = out if gradient else (out, None)
        iso = Z[gsz:][sl_sp][st.ref_idx]
        ids = unit_values[su0[i]:su0[i + 1] + 1]
        block = activate(Z, iso, ids, options.sigmoid_slope)
        if rel[i] == FAULT:
            values_everywhere[i] = block - block.min()
        else:
            values_everywhere[i] = block
        results.append(StackResult(w, Z, G, iso, block, rel[i], cond=cond))
    # ---- masks (ERODE: above the stack's lowest surface; FAULT: none; BASEMENT: everything)
    masks = np.zeros((n_st, xyz.shape[0]), bool)
    for i, r in enumerate(results):
        if rel[i] == ERODE:
            masks[i] = r.Z > r.isovalues.min()
        elif rel[i] == ONLAP:
            nxt = results[i + 1]
            masks[i] = nxt.Z > nxt.isovalues.max()
        elif rel[i] == FAULT:
            masks[i] = False
        else:
            masks[i] = True
    for i in range(n_st - 2, -1, -1):           # chained onlaps
        if rel[i] == ONLAP and rel[i + 1] == ONLAP:
            masks[i] &= masks[i + 1]
    free = np.ones(xyz.shape[0], bool)
    final = np.zeros(xyz.shape[0])
    faults = np.zeros(xyz.shape[0])
    for i, r in enumerate(results):
        r.mask = masks[i]
        r.squeezed_mask = masks[i] & free
        free = free & ~masks[i]
        if rel[i] == FAULT:
            faults += r.values_block
        else:
            final += r.values_block * r.squeezed_mask
    return FieldsResult(results, final, faults, gsz)


# ----------------------------------------------------------------------------------------------
# octree
# ----------------------------------------------------------------------------------------------
_SX = np.array([-1, -1, -1, -1, 1, 1, 1, 1], float)
_SY = np.array([-1, -1, 1, 1, -1, -1, 1, 1], float)
_SZ = np.array([-1, 1, -1, 1, -1, 1, -1, 1], float)
_S8 = np.stack([_SX, _SY, _SZ], axis=1)            # (8, 3) corner / child sign pattern


def regular_grid_centers(extent, shape, shift: float = None):
    """Cell centres (x slowest, z fastest) + GRID_SHIFT, and the cell size."""
    e = np.asarray(extent, float)
    s = np.asarray(shape, int)
    d = np.array([(e[1] - e[0]) / s[0], (e[3] - e[2]) / s[1], (e[5] - e[4]) / s[2]])
    ax = [np.linspace(e[2 * k] + d[k] / 2, e[2 * k + 1] - d[k] / 2, int(s[k])) for k in range(3)]
    g = np.meshgrid(*ax, indexing="ij")
    return np.vstack([c.ravel() for c in g]).T + (GRID_SHIFT if shift is None else shift), d


def voxel_corners(centers: np.ndarray, dxdydz: np.ndarray) -> np.ndarray:
    """8 corners per voxel, voxel-major; sign pattern x:----++++ y:--++--++ z:-+-+-+-+."""
    return (centers[:, None, :] + _S8[None, :, :] * (dxdydz / 2)[None, None, :]).reshape(-1, 3)


def voxel_children(centers: np.ndarray, dxdydz: np.ndarray) -> np.ndarray:
    return (centers[:, None, :] + _S8[None, :, :] * (dxdydz / 4)[None, None, :]).reshape(-1, 3)


def mark_voxels_by_corners(ids_corners: np.ndarray) -> np.ndarray:
    """Refine a voxel when its 8 corner ids differ (engine stage "octrees_topology"; the rule
    that reproduces the leaf lists behind test/test_model_types/*.approved.txt)."""
    u = ids_corners.reshape(-1, 8)
    return (u != u[:, :1]).any(axis=1)


def mark_voxels_by_stats(Z_centers: np.ndarray, isovalues: np.ndarray, selected_by_corners: np.ndarray,
                         n_std: float) -> np.ndarray:
    """Voxels whose centre value is within mean + n_std*std of the distance-to-nearest-isovalue
    statistics of the corner-selected voxels."""
    d = np.abs(Z_centers[:, None] - np.asarray(isovalues)[None, :]).min(axis=1)
    if not selected_by_corners.any():
        return np.zeros_like(selected_by_corners)
    ref = d[selected_by_corners]
    return np.abs(d - ref.mean()) < n_std * ref.std()


@dataclasses.dataclass
class OracleLevel:
    centers: np.ndarray
    dxdydz: np.ndarray
    fields: FieldsResult
    corners: Optional[np.ndarray] = None
    fields_corners: Optional[FieldsResult] = None
    selected: Optional[np.ndarray] = None


def interpolate_n_octree_levels(interp_input, options, descriptor) -> List[OracleLevel]:
    """Level loop (docs/developers_notes/dev_log/log_2024-05.md:7-8, log_2024-06.md:27-39); root grid = regular grid of
    the octree base resolution (gempy/core/data/grid.py:127-151, _engine_factory.py:88-96); levels below
    evaluation_options.octree_min_level are refined everywhere."""
    eo = options.evaluation_options
    og = interp_input.grid.octree_grid
    centers, d = regular_grid_centers(og.orthogonal_extent, og.regular_grid_shape)
    extra = [g.values for n, g in interp_input.grid.parts() if n != "octree_grid"]
    n_levels = eo.number_octree_levels
    cache = [None] * descriptor.stack_structure.n_stacks
    if interp_input.weights:
        for i, w in enumerate(interp_input.weights):
            cache[i] = np.asarray(w, float) if w is not None and len(w) else None
    levels: List[OracleLevel] = []
    for lvl in range(n_levels):
        pts = np.vstack([centers] + (extra if lvl == 0 else []))
        f_c = interpolate_all_fields(interp_input, options, descriptor, pts, cache)
        level = OracleLevel(centers, d.copy(), f_c)
        need_corners = (lvl < n_levels - 1) or eo.mesh_extraction
        if need_corners:
            level.corners = voxel_corners(centers, d)
            level.fields_corners = interpolate_all_fields(interp_input, options, descriptor, level.corners, cache,
                                                          gradient=False)
        levels.append(level)
        if lvl == n_levels - 1:
            break
        nv = centers.shape[0]
        sel = mark_voxels_by_corners(level.fields_corners.litho_faults_ids)
        if lvl < eo.octree_min_level:
            sel = np.ones(nv, bool)
        elif USE_STATS_REFINEMENT and eo.octree_error_threshold > 0:
            extra_sel = np.zeros(nv, bool)
            for st in f_c.stacks:
                extra_sel |= mark_voxels_by_stats(st.Z[:nv], st.isovalues, sel, eo.octree_error_threshold)
            sel = sel | extra_sel
        level.selected = sel
        centers = voxel_children(centers[sel], d)
        d = d / 2
    return levels


# ----------------------------------------------------------------------------------------------
# dual contouring
# ----------------------------------------------------------------------------------------------
_EDGE_A = np.array([0, 1, 2, 3, 0, 1, 4, 5, 0, 2, 4, 6])     # lower corner of the 12 edges (x x x x y y y y z z z z)
_EDGE_B = np.array([4, 5, 6, 7, 2, 3, 6, 7, 1, 3, 5, 7])     # upper corner
_EDGE_AXIS = np.array([0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2])


def edge_intersections(corners_xyz: np.ndarray, Z_corners: np.ndarray, iso: float):
    """Per voxel, per edge: crossing of the isovalue by linear interpolation.
    Returns valid (nv,12) bool and xyz (nv,12,3) (zeros where invalid)."""
    c = corners_xyz.reshape(-1, 8, 3)
    z = Z_corners.reshape(-1, 8)
    za, zb = z[:, _EDGE_A], z[:, _EDGE_B]
    with np.errstate(divide="ignore", invalid="ignore"):
        wgt = (iso - zb) / (za - zb)            # weight towards corner A measured from B
    valid = (wgt > 0) & (wgt < 1)
    pa, pb = c[:, _EDGE_A, :], c[:, _EDGE_B, :]
    xyz = pb + (pa - pb) * np.where(valid, wgt, 0.0)[:, :, None]
    xyz = np.where(valid[:, :, None], xyz, 0.0)
    return valid, xyz


def dual_contour_vertices(valid: np.ndarray, xyz: np.ndarray, grads: np.ndarray, bias_strength: float = 1.0):
    # engine stage "dual_contouring" (SURVEY.md 8a2 row 4c); consumers: gempy/core/data/geo_model.py:110-121.  Unpinned.
    """QEF per voxel: 12 edge planes (normal = raw gradient at the crossing) + 3 axis planes through the
    mass point.  Coordinates that are within 1e-8 of zero are ignored by the mass point (upstream quirk)."""
    vv = valid.any(axis=1)
    e_xyz = xyz[vv]
    e_n = np.where(valid[vv][:, :, None], grads[vv], 0.0)
    use = valid[vv][:, :, None] & ~np.isclose(e_xyz, 0.0)
    cnt = use.sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        mass = np.where(use, e_xyz, 0.0).sum(axis=1) / cnt
    A = np.concatenate([e_n, np.broadcast_to(np.eye(3) * bias_strength, (e_n.shape[0], 3, 3))], axis=1)
    P = np.concatenate([e_xyz, np.repeat(mass[:, None, :], 3, axis=1)], axis=1)
    b = (A * P).sum(axis=2)
    AtA = np.einsum("vki,vkj->vij", A, A)
    Atb = np.einsum("vki,vk->vi", A, b)
    verts = np.linalg.solve(AtA, Atb[:, :, None])[:, :, 0]
    return verts, vv


def dual_contour_triangles(valid: np.ndarray, voxel_ijk: np.ndarray):
    """For every crossed edge shared by four existing surface voxels emit a quad as two triangles.
    voxel_ijk are integer lattice coordinates of the voxels; vertex index = rank among valid voxels."""
    vv = valid.any(axis=1)
    ijk = voxel_ijk[vv]
    lut = {tuple(k): n for n, k in enumerate(ijk.tolist())}
    val = valid[vv]
    tris = []
    # edge e of voxel v (axis ax, at the voxel's low/high side in the two other axes) is shared with the
    # three voxels offset in those two axes; emit the quad from the voxel for which the edge is its
    # "high-high" edge so each edge is emitted once.
    hh_edge = {0: 3, 1: 7, 2: 11}       # x-edge with y+,z+ ; y-edge with x+,z+ ; z-edge with x+,y+
    others = {0: (1, 2), 1: (0, 2), 2: (0, 1)}
    for ax in range(3):
        e = hh_edge[ax]
        u, v = others[ax]
        for n in np.nonzero(val[:, e])[0]:
            k = ijk[n]
            ku = k.copy(); ku[u] += 1
            kv = k.copy(); kv[v] += 1
            kuv = ku.copy(); kuv[v] += 1
            a, b, c = lut.get(tuple(ku)), lut.get(tuple(kv)), lut.get(tuple(kuv))
            if a is None or b is None or c is None:
                continue
            tris.append((n, a, c))
            tris.append((n, c, b))
    return np.asarray(tris, dtype=np.int64).reshape(-1, 3)


@dataclasses.dataclass
class OracleMesh:
    vertices: np.ndarray
    edges: np.ndarray
    stack: int
    surface: int


def lattice_ijk(centers, extent, dxdydz):
    e = np.asarray(extent, float)
    return np.rint((centers - e[[0, 2, 4]]) / dxdydz - 0.5).astype(np.int64)


def dual_contouring(interp_input, options, descriptor, levels: List[OracleLevel], weights_cache) -> List[OracleMesh]:
    eo = options.evaluation_options
    lvl = levels[min(eo.number_octree_levels_surface, len(levels)) - 1]
    ko = options.kernel_options
    sp0, or0, su0 = _stack_slices(descriptor)
    ss = descriptor.stack_structure
    extent = interp_input.grid.octree_grid.orthogonal_extent
    ijk = lattice_ijk(lvl.centers, extent, lvl.dxdydz)
    meshes = []
    nvox = lvl.centers.shape[0]
    for i, st_res in enumerate(lvl.fields_corners.stacks):
        Zc = st_res.Z[:nvox * 8]
        # masking: INTERSECT-like -- only voxels where this stack owns at least one corner
        own = st_res.squeezed_mask[:nvox * 8].reshape(-1, 8).any(axis=1) if st_res.relation != FAULT else np.ones(nvox, bool)
        for s, iso in enumerate(st_res.isovalues):
            valid, xyz = edge_intersections(lvl.corners, Zc, iso)
            valid &= own[:, None]
            xyz = np.where(valid[:, :, None], xyz, 0.0)
            pts = xyz[valid]
            # gradient of this stack's field at the crossings
            g = _stack_gradient_at(interp_input, options, descriptor, i, pts, weights_cache, levels)
            grads = np.zeros_like(xyz)
            grads[valid] = g
            verts, vv = dual_contour_vertices(valid, xyz, grads)
            tris = dual_contour_triangles(valid, ijk)
            meshes.append(OracleMesh(verts, tris, i, s))
    return meshes


def _stack_gradient_at(interp_input, options, descriptor, i, pts, weights_cache, levels):
    if pts.shape[0] == 0:
        return np.zeros((0, 3))
    f = interpolate_all_fields(interp_input, options, descriptor, pts, weights_cache, gradient=True)
    return f.stacks[i].G[:pts.shape[0]]


# ----------------------------------------------------------------------------------------------
# forward gravity
# ----------------------------------------------------------------------------------------------
def forward_gravity(interp_input, options, descriptor, tz, densities) -> np.ndarray:
    """gravity[c] = sum_k tz[k] * density[lith id at (centre c + kernel voxel k)]
    (known answer: test/test_modules/test_geophysics/test_gravity.py:89)."""
    g = interp_input.grid.geophysics_grid
    f = interpolate_all_fields(interp_input, options, descriptor, g.values, gradient=False)
    ids = np.clip(f.lith_ids.astype(int), 1, len(densities))
    dens = np.asarray(densities, float)[ids - 1].reshape(g.centers.shape[0], -1)
    return (dens * np.asarray(tz, float)[None, :]).sum(axis=1)


# ----------------------------------------------------------------------------------------------
# condition number and its derivative with respect to the surface-point nuggets
# ----------------------------------------------------------------------------------------------
def condition_number_and_gradient(interp_input, options, descriptor, i: int, fd: bool = False):
    """2-norm condition number of stack i's matrix (no fault columns: the reference optimises one group in isolation,
    gempy/modules/optimize_nuggets/_optimizer.py:48-52) and d cond / d nugget of the stack's surface points.
    Gradient from the extreme eigenpairs of the symmetric matrix; ``fd=True`` returns central finite differences
    instead (the check of the analytic form).  Parity unpinned (engine definition absent from the reference tree)."""
    ko = options.kernel_options
    sp0, or0, su0 = _stack_slices(descriptor)
    sl_sp, sl_or, sl_su = slice(sp0[i], sp0[i + 1]), slice(or0[i], or0[i + 1]), slice(su0[i], su0[i + 1])
    nps = np.asarray(descriptor.tensors_structure.number_of_points_per_surface[sl_su], dtype=int)

    def matrix(nug):
        st = prepare_stack(interp_input.surface_points.sp_coords[sl_sp], nug, nps,
                           interp_input.orientations.dip_positions[sl_or], interp_input.orientations.dip_gradients[sl_or],
                           interp_input.orientations.nugget_effect_grad[sl_or])
        return assemble_covariance(st, ko)

    nug0 = np.asarray(interp_input.surface_points.nugget_effect_scalar[sl_sp], dtype=float).copy()
    A = matrix(nug0)
    if fd:
        g = np.zeros_like(nug0)
        for k in range(nug0.size):
            h = 1e-6 * max(nug0[k], 1e-4)
            up, dn = nug0.copy(), nug0.copy()
            up[k] += h
            dn[k] -= h
            g[k] = (np.linalg.cond(matrix(up)) - np.linalg.cond(matrix(dn))) / (2 * h)
        return float(np.linalg.cond(A)), g
    lam, Q = np.linalg.eigh(A)
    iM, im = int(np.argmax(np.abs(lam))), int(np.argmin(np.abs(lam)))
    n_ori = interp_input.orientations.dip_positions[sl_or].shape[0]
    starts = np.concatenate([[0], np.cumsum(nps)[:-1]])
    is_ref = np.zeros(nug0.size, bool)
    is_ref[starts] = True
    rest_idx = np.nonzero(~is_ref)[0]
    ref_idx = np.repeat(starts, nps - 1)
    rows = slice(3 * n_ori, 3 * n_ori + rest_idx.size)
    g_diag = (np.sign(lam[iM]) * Q[rows, iM] ** 2 * abs(lam[im]) - abs(lam[iM]) * np.sign(lam[im]) * Q[rows, im] ** 2) / lam[im] ** 2
    g = np.zeros_like(nug0)
    np.add.at(g, rest_idx, 0.5 * ko.c_o * g_diag)
    np.add.at(g, ref_idx, 0.5 * ko.c_o * g_diag)
    return float(abs(lam[iM]) / abs(lam[im])), g


# ----------------------------------------------------------------------------------------------
# marching cubes on the dense grid
# ----------------------------------------------------------------------------------------------
# The reference's dense-grid mesher (gempy/modules/mesh_extranction/marching_cubes.py:13-101) hands each stack's
# scalar field on the dense grid, the element's isovalue and the stack's squeezed mask to
# skimage.measure.marching_cubes(method="lewiner", allow_degenerate=False) -- a third-party routine absent from this
# image (scikit-image; gempy's optional dependency).  Restated here from its published behaviour:
#   * one vertex per lattice edge whose end values straddle the level, shared by the cubes around the edge;
#   * a cube is processed only if the mask is set at its far corner (i+1, j+1, k+1)  [pinned, see below];
#   * vertex = linear interpolation along the edge, in index units * spacing, then + (extent minima)
#     (marching_cubes.py:82-95: the half-cell offset of the cell centres is NOT added -- kept as the reference does).
# PIN: test/test_modules/test_marching_cubes.py:44-47 -- COMBINATION on a dense 40 x 20 x 20 grid gives exactly
# 600 / 860 / 1256 / 1680 vertices for fault / rock3 / rock2 / rock1; this restatement reproduces all four
# (tests/test_oracle.py::test_marching_cubes_vertex_counts), which also pins the dense-grid fields, the isovalues
# and the squeezed ERODE masks.  The TRIANGLE table is not pinned (Lewiner resolves ambiguous faces by an interior
# test; here ambiguous faces always separate the above-level corners): "parity unpinned" for faces.
MC_EDGE_CORNERS = ((0, 4), (1, 5), (2, 6), (3, 7), (0, 2), (1, 3), (4, 6), (5, 7), (0, 1), (2, 3), (4, 5), (6, 7))


def _mc_faces():
    """Six faces as corner cycles, counter-clockwise seen from outside.  corner id = 4*x + 2*y + z."""
    faces = []
    for axis in range(3):
        for side in (0, 1):
            u, v = [(1, 2), (2, 0), (0, 1)][axis]
            if side == 0:
                u, v = v, u                     # u x v must be the outward normal
            cyc = []
            for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = [0, 0, 0]
                c[axis], c[u], c[v] = side, a, b
                cyc.append(4 * c[0] + 2 * c[1] + c[2])
            faces.append(cyc)
    return faces


def marching_cubes_table():
    """256 cases -> list of triangles (edge-id triples).  On every face the crossings are joined so that each run of
    above-level corners is cut off by its own segment (ambiguous faces separate the above-level corners); segments
    run from the edge where the counter-clockwise walk leaves the run to the edge where it entered it, the closed
    loops are fanned from their lowest edge id, and the winding makes normals point towards lower values."""
    edge_of = {}
    for e, (a, b) in enumerate(MC_EDGE_CORNERS):
        edge_of[(a, b)] = e
        edge_of[(b, a)] = e
    faces = _mc_faces()
    table = []
    for case in range(256):
        inside = [(case >> c) & 1 for c in range(8)]
        nxt = {}
        for cyc in faces:
            for q in range(4):
                # a run of inside corners starts at cyc[q] when the previous corner is outside
                if inside[cyc[q]] and not inside[cyc[q - 1]]:
                    entry = edge_of[(cyc[q - 1], cyc[q])]
                    r = q
                    while inside[cyc[(r + 1) % 4]]:
                        r += 1
                    leave = edge_of[(cyc[r % 4], cyc[(r + 1) % 4])]
                    nxt[leave] = entry
        tris, seen = [], set()
        for e0 in sorted(nxt):
            if e0 in seen:
                continue
            loop, e = [], e0
            while e not in seen:
                seen.add(e)
                loop.append(e)
                e = nxt[e]
            for t in range(1, len(loop) - 1):
                tris.append((loop[0], loop[t + 1], loop[t]))
        table.append(tris)
    return table


_MC_TABLE = None


def marching_cubes(Z, shape, level, mask=None, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)):
    """-> vertices (V,3) float64, triangles (T,3) int64.  Vertex order: owner lattice point (x slowest, z fastest),
    then edge axis x, y, z; triangle order: cube index, then table order."""
    global _MC_TABLE
    if _MC_TABLE is None:
        _MC_TABLE = marching_cubes_table()
    nx, ny, nz = (int(v) for v in shape)
    Z = np.asarray(Z, float).reshape(nx, ny, nz)
    ins = Z > level
    ok = np.ones((nx - 1, ny - 1, nz - 1), bool) if mask is None else \
        np.asarray(mask).reshape(nx, ny, nz)[1:, 1:, 1:].astype(bool)
    # edges needed by at least one processed cube
    need = [np.zeros((nx - 1, ny, nz), bool), np.zeros((nx, ny - 1, nz), bool), np.zeros((nx, ny, nz - 1), bool)]
    for a in (0, 1):
        for b in (0, 1):
            need[0][:, a:ny - 1 + a, b:nz - 1 + b] |= ok
            need[1][a:nx - 1 + a, :, b:nz - 1 + b] |= ok
            need[2][a:nx - 1 + a, b:ny - 1 + b, :] |= ok
    present = np.zeros((nx, ny, nz, 3), bool)
    present[:-1, :, :, 0] = (ins[1:] != ins[:-1]) & need[0]
    present[:, :-1, :, 1] = (ins[:, 1:] != ins[:, :-1]) & need[1]
    present[:, :, :-1, 2] = (ins[:, :, 1:] != ins[:, :, :-1]) & need[2]
    flat = present.reshape(-1)
    index = np.cumsum(flat) - 1                          # vertex id of (point, axis)
    pts, axes = np.nonzero(present.reshape(-1, 3))
    i, rem = np.divmod(pts, ny * nz)
    j, k = np.divmod(rem, nz)
    ijk = np.stack([i, j, k], axis=1)
    nb = ijk.copy()
    nb[np.arange(len(axes)), axes] += 1
    z0 = Z[i, j, k]
    z1 = Z[nb[:, 0], nb[:, 1], nb[:, 2]]
    t = (level - z0) / (z1 - z0)
    pos = ijk.astype(float)
    pos[np.arange(len(axes)), axes] += t
    verts = pos * np.asarray(spacing, float)[None, :] + np.asarray(origin, float)[None, :]
    # triangles
    case = np.zeros((nx - 1, ny - 1, nz - 1), int)
    for c in range(8):
        cx, cy, cz = c >> 2, (c >> 1) & 1, c & 1
        case |= ins[cx:nx - 1 + cx, cy:ny - 1 + cy, cz:nz - 1 + cz].astype(int) << c
    index3 = index.reshape(nx, ny, nz, 3)
    tris = []
    for (ci, cj, ck) in np.argwhere(ok & (case > 0) & (case < 255)):
        for tri in _MC_TABLE[case[ci, cj, ck]]:
            row = []
            for e in tri:
                c0 = MC_EDGE_CORNERS[e][0]
                row.append(index3[ci + (c0 >> 2), cj + ((c0 >> 1) & 1), ck + (c0 & 1), e // 4])
            tris.append(row)
    return verts, np.asarray(tris, dtype=np.int64).reshape(-1, 3)


def marching_cubes_meshes(fields: "FieldsResult", descriptor, dense_shape, dense_slice: slice, real_extent):
    """What set_meshes_with_marching_cubes leaves on the structural elements (marching_cubes.py:37-55): per stack,
    per surface, (vertices, triangles) in real coordinates; faults use no mask, other stacks their squeezed mask."""
    shape = np.asarray(dense_shape, int)
    ext = np.asarray(real_extent, float)
    spacing = (ext[1::2] - ext[0::2]) / shape
    out = []
    for st in fields.stacks:
        Z = st.Z[dense_slice]
        mask = None if st.relation == FAULT else st.squeezed_mask[dense_slice]
        for iso in st.isovalues:
            out.append(marching_cubes(Z, shape, iso, mask, spacing, ext[0::2]))
    return out


# ----------------------------------------------------------------------------------------------
# entry
# ----------------------------------------------------------------------------------------------
def fill_regular_from_octree(levels_host, base_shape, key) -> np.ndarray:
    """Dense array at the finest octree resolution: level-0 values upsampled, refined voxels overwritten by their
    children (the engine's octree -> regular fill behind RawArraysSolution.lith_block of an octree model; consumers
    gempy/API/gp2_gp3_compatibility/gp3_to_gp2_output.py:55-88).  levels_host: per level a dict with the values (read
    through `key`) and "selected" = the refine mask of the level (None on the last)."""
    shape = np.asarray(base_shape, dtype=int)
    dense = np.asarray(key(levels_host[0])).reshape(shape)
    ijk = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), axis=-1).reshape(-1, 3)
    for lvl in range(1, len(levels_host)):
        sel = np.asarray(levels_host[lvl - 1]["selected"], dtype=bool)
        dense = dense.repeat(2, axis=0).repeat(2, axis=1).repeat(2, axis=2)
        parents = ijk[sel]
        off = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)])
        ijk = (parents[:, None, :] * 2 + off[None, :, :]).reshape(-1, 3)
        dense[ijk[:, 0], ijk[:, 1], ijk[:, 2]] = key(levels_host[lvl])
    return dense.ravel()


@dataclasses.dataclass
class OracleSolutions:
    levels: List[OracleLevel]
    meshes: Optional[List[OracleMesh]]
    weights: List[np.ndarray]


def compute_model(interpolation_input, options, data_descriptor, geophysics_input=None) -> OracleSolutions:
    levels = interpolate_n_octree_levels(interpolation_input, options, data_descriptor)
    weights = [s.weights for s in levels[0].fields.stacks]
    meshes = None
    if options.evaluation_options.mesh_extraction:
        meshes = dual_contouring(interpolation_input, options, data_descriptor, levels, list(weights))
    return OracleSolutions(levels, meshes, weights)
