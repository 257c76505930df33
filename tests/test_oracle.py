"""The oracle against the reference's own golden vectors and known answers (CPU only).

Goldens: tests/golden/approved_scalar_fields.json, extracted by tests/golden/make_fixtures.py from
/root/reference/test/test_model_types/*.approved.txt (test_example_models_I.py:19-88)."""
import json
import os

import numpy as np
import pytest

from gempy_b200 import examples as ex
from oracle import gempy_oracle as orc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "approved_scalar_fields.json")))


def _verify_scalar_field(levels, with_corners):
    """Restates _verify_scalar_field (test_example_models_I.py:19-23): stack 0 of the last octree level,
    every len//50-th value.  When the last level is also the dual-contouring level the engine's array holds
    centres ++ corners."""
    last = levels[-1]
    n1 = last.centers.shape[0]
    z = last.fields.stacks[0].Z[:n1]
    if with_corners:
        z = np.concatenate([z, last.fields_corners.stacks[0].Z[:8 * n1]])
    return z[::int(len(z) / 50)]


@pytest.mark.parametrize("key,build,with_corners", [
    ("anticline", ex.anticline, False),       # refinement 5, surface level 4 -> last level has no corners
    ("fault", ex.one_fault, False),           # refinement 6
    ("combination", ex.combination, True),    # refinement 4 == number_octree_levels_surface
])
def test_approved_scalar_fields(key, build, with_corners):
    m = build()
    ii, opt, desc = m.args()
    opt.evaluation_options.mesh_extraction = with_corners
    levels = orc.interpolate_n_octree_levels(ii, opt, desc)
    got = _verify_scalar_field(levels, with_corners)
    want = np.array(GOLD[key])
    assert got.shape == want.shape == (51,)
    # the reference's comparator is allclose(rtol=1e-5, atol=1e-5) (test/verify_helper.py:70-101);
    # the approved file prints 8 significant digits, so 5e-8 is the tightest meaningful bound
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-8)


def test_greenstone_isovalues_stored_by_the_engine():
    """examples/data/gempy_models/Greenstone.gempy keeps, in its header, the scalar field at the interfaces the real
    engine computed (16 significant digits).  Three series, 70 surface points, 41 orientations."""
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "greenstone_isovalues.json")))
    m = ex.greenstone()
    np.testing.assert_allclose(m.transform.scale, [1.0296222316032248e-05] * 3, rtol=1e-15)
    ii, opt, desc = m.args()
    f = orc.interpolate_all_fields(ii, opt, desc, np.zeros((1, 3)))
    got = {}
    k = 0
    for st in f.stacks:
        for iso in st.isovalues:
            got[m.element_names[k]] = iso
            k += 1
    assert set(got) == set(want)
    for name, v in want.items():
        assert abs(got[name] - v) < 1e-11, (name, got[name], v)          # measured: 2e-12 / 3e-13


def test_gravity_known_answer():
    """test/test_modules/test_geophysics/test_gravity.py:89: np.testing.assert_almost_equal(gravity, [-1624.1714], 4)."""
    m, geo = ex.two_layers_gravity(resolution=(10, 1, 10))
    ii, opt, desc = m.args()
    g = orc.forward_gravity(ii, opt, desc, geo.tz, geo.densities)
    np.testing.assert_almost_equal(g, np.array([-1624.1714]), decimal=4)
    # the stored isovalues of the same model (2-layers.approved.txt): 0.09000000000000002 and -0.2483333333333333
    f = orc.interpolate_all_fields(ii, opt, desc, np.zeros((1, 3)))
    np.testing.assert_allclose(f.stacks[0].isovalues, [0.09000000000000002, -0.2483333333333333], rtol=0, atol=1e-12)


def test_custom_grid_known_answer():
    """test/test_modules/test_grids/test_custom_grid.py:24-47."""
    xyz = np.array([[0, 0, 0], [1000, 0, 0], [0, 1000, 0], [1000, 1000, 0],
                    [0, 0, 1000], [1000, 0, 1000], [0, 1000, 1000], [1000, 1000, 1000]], dtype=float)
    m = ex.anticline(custom_xyz=xyz)
    ii, opt, desc = m.args()
    f = orc.interpolate_all_fields(ii, opt, desc, ii.grid.custom_grid.values)
    np.testing.assert_array_equal(f.lith_ids, np.array([3., 3., 3., 3., 1., 1., 1., 1.]))


def test_horizontal_plane_and_transform():
    """Golden JSON pins the input transform of HORIZONTAL_STRAT (position -500, scale 6.25e-4); the kriged
    field of two horizontal layers is the exact plane Z = gi_res * z'."""
    m = ex.horizontal_strat()
    np.testing.assert_allclose(m.transform.position, [-500, -500, -500])
    np.testing.assert_allclose(m.transform.scale, [0.000625] * 3)
    ii, opt, desc = m.args()
    xyz = ii.grid.dense_grid.values
    f = orc.interpolate_all_fields(ii, opt, desc, xyz, gradient=True)
    st = f.stacks[0]
    np.testing.assert_allclose(st.Z[:xyz.shape[0]], 2.0 * xyz[:, 2], atol=1e-12)
    np.testing.assert_allclose(st.G[:xyz.shape[0]], np.tile([0, 0, 1.0], (xyz.shape[0], 1)), atol=1e-12)
    ids, counts = np.unique(f.lith_ids, return_counts=True)
    assert ids.tolist() == [1., 2., 3.] and counts.tolist() == [5000, 2500, 5000]


def test_scalar_field_matrix_shape_contract():
    """test/test_modules/test_outliers.py:51: (n_stacks, n_dense_points)."""
    m = ex.combination()
    ii, opt, desc = m.args()
    c, _ = orc.regular_grid_centers(ii.grid.octree_grid.orthogonal_extent, [10, 5, 5])
    f = orc.interpolate_all_fields(ii, opt, desc, c)
    mat = np.stack([s.Z[:f.grid_size] for s in f.stacks])
    assert mat.shape == (3, 250)


def test_interpolant_honours_data():
    """Mathematical self-check: Z(rest_i) = Z(ref_i) up to the nugget, gradient at the
    orientations ~ G (engine gradient convention)."""
    m = ex.synthetic_stress(n_sp_per_surface=40, n_surfaces=3, n_ori=30, resolution=(4, 4, 4))
    ii, opt, desc = m.args()
    f = orc.interpolate_all_fields(ii, opt, desc, ii.orientations.dip_positions, gradient=True)
    st = f.stacks[0]
    n_o = ii.orientations.n_items
    zsp = st.Z[n_o:]
    starts = desc.tensors_structure.reference_sp_position
    n = desc.tensors_structure.number_of_points_per_surface
    for s0, k in zip(starts, n):
        assert np.abs(zsp[s0:s0 + k] - zsp[s0]).max() < 5e-3
    g = st.G[:n_o]
    assert np.abs(g - ii.orientations.dip_gradients).max() < 5e-2


def test_gradient_matches_finite_difference():
    """Engine-convention gradient = (1/gi_res) dZ/dx away from the data (regulariser negligible there)."""
    m = ex.anticline()
    ii, opt, desc = m.args()
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.2, 0.2, size=(20, 3))
    h = 1e-6
    f0 = orc.interpolate_all_fields(ii, opt, desc, x, gradient=True).stacks[0]
    for a in range(3):
        dx = np.zeros(3); dx[a] = h
        zp = orc.interpolate_all_fields(ii, opt, desc, x + dx).stacks[0].Z[:20]
        zm = orc.interpolate_all_fields(ii, opt, desc, x - dx).stacks[0].Z[:20]
        fd = (zp - zm) / (2 * h) / opt.kernel_options.gi_res
        np.testing.assert_allclose(f0.G[:20, a], fd, rtol=2e-3, atol=2e-3)


def test_dual_contouring_vertices_near_surface():
    m = ex.anticline(refinement=4)
    ii, opt, desc = m.args()
    sol = orc.compute_model(ii, opt, desc)
    assert sol.meshes is not None and len(sol.meshes) == 2
    for mesh in sol.meshes:
        assert mesh.vertices.shape[0] > 50 and mesh.edges.shape[0] > 50
        f = orc.interpolate_all_fields(ii, opt, desc, mesh.vertices)
        z = f.stacks[0].Z[:mesh.vertices.shape[0]]
        iso = f.stacks[0].isovalues[mesh.surface]
        assert np.abs(z - iso).max() < 0.03        # vertices sit on the isosurface to a fraction of a voxel
        assert mesh.edges.max() < mesh.vertices.shape[0]


def test_marching_cubes_vertex_counts():
    """test/test_modules/test_marching_cubes.py:13-47: COMBINATION on a dense 40 x 20 x 20 grid; the reference pins
    600 / 860 / 1256 / 1680 vertices for fault / rock3 / rock2 / rock1.  Reproducing all four pins the dense-grid
    fields (fault drift included), the isovalues, the squeezed ERODE masks and the far-corner mask rule."""
    m = ex.combination(refinement=None, resolution=(40, 20, 20))
    ii, opt, desc = m.args()
    g = ii.grid.dense_grid
    f = orc.interpolate_all_fields(ii, opt, desc, g.values + orc.GRID_SHIFT)
    meshes = orc.marching_cubes_meshes(f, desc, g.regular_grid_shape, slice(0, g.n_points), m.extent)
    assert [v.shape[0] for v, _ in meshes] == [600, 860, 1256, 1680]
    # vertices come back in real coordinates inside the model extent (marching_cubes.py:92-95)
    for v, t in meshes:
        assert (v.min(0) >= -1e-9).all() and (v.max(0) <= np.array([2500, 1000, 1000]) + 1e-9).all()
        assert t.min() == 0 and t.max() == v.shape[0] - 1


def test_marching_cubes_table_is_watertight():
    """Closed surface of a smooth blob: every directed edge is matched by its reverse, Euler characteristic 2,
    normals towards lower values; the 256-case table never exceeds 5 triangles per cube."""
    table = orc.marching_cubes_table()
    assert max(len(t) for t in table) == 5 and len(table[0]) == 0 and len(table[255]) == 0
    for case in range(256):                      # complementary cases cut the same edges
        assert sorted({e for t in table[case] for e in t}) == sorted({e for t in table[255 - case] for e in t})
    n = 20
    ax = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    F = 0.6 - np.sqrt(X ** 2 + 1.3 * Y ** 2 + 0.8 * Z ** 2) + 0.15 * np.sin(5 * X) * np.cos(4 * Y)
    v, t = orc.marching_cubes(F, (n, n, n), 0.0)
    edges = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    directed = set(map(tuple, edges.tolist()))
    assert len(directed) == edges.shape[0]
    assert all((b, a) in directed for a, b in directed)
    assert v.shape[0] - len(directed) // 2 + t.shape[0] == 2
    tri = v[t]
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert (np.einsum("ij,ij->i", nrm, tri.mean(1) - (n - 1) / 2) > 0).all()


def test_condition_number_gradient_and_nugget_optimiser():
    """gempy/modules/optimize_nuggets (_optimizer.py:9-68, _ops.py:6-94) restated: the analytic derivative of the
    condition number with respect to the surface-point nuggets agrees with central finite differences, and the Adam
    loop (lr 0.01, top-1 % gradient masking, clamp at 1e-7, the reference's convergence rule) lowers the condition
    number of an ill-conditioned stack by orders of magnitude."""
    from gempy_b200.engine.nuggets import gradient_masking, has_converged, optimize_nuggets
    m = ex.anticline()
    ii, opt, desc = m.args()
    c, g = orc.condition_number_and_gradient(ii, opt, desc, 0)
    c_fd, g_fd = orc.condition_number_and_gradient(ii, opt, desc, 0, fd=True)
    assert abs(c - c_fd) < 1e-6 * c
    assert np.abs(g - g_fd).max() < 1e-4 * np.abs(g).max()
    # helpers follow the reference's rules
    gm = gradient_masking(np.array([1.0, -5.0, 3.0, 0.5] * 50), focus=0.01)          # int(200 * 0.01) = 2 entries survive
    assert np.count_nonzero(gm) == 2 and set(np.unique(gm)) == {-5.0, 0.0}
    assert gradient_masking(np.ones(27), focus=0.01).sum() == 0                       # int(27 * 0.01) = 0, as torch.topk(k=0)
    assert has_converged(9e4, 1e9) and not has_converged(2e5, 1e9, epoch=3)
    assert has_converged(2e5, 2.001e5, epoch=11) and not has_converged(2e5, 3e5, epoch=11)
    big = ex.synthetic_stress(n_sp_per_surface=60, n_surfaces=4, n_ori=40, resolution=(4, 4, 4))
    ii, opt, desc = big.args()
    hist = optimize_nuggets(ii, opt, desc, max_epochs=40, convergence_criteria=1e3,
                            cond_and_grad=lambda i: orc.condition_number_and_gradient(ii, opt, desc, i))
    assert hist[0][0] > 5e6 and hist[0][-1] < 1e5
    nug = ii.surface_points.nugget_effect_scalar
    assert (nug >= 1e-7).all() and 0 < (nug != 2e-5).sum() <= 40
    assert opt.kernel_options.condition_number == hist[0][-1]


def test_oracle_invariances():
    """Properties universal co-kriging must have whatever the conventions: the gradient field is invariant under a
    common translation of data and evaluation points (degree-1 drift), the field moves by a constant only, and the
    result does not depend on the order of the non-reference points of a surface or of the orientations."""
    m = ex.anticline()
    ii, opt, desc = m.args()
    ko = opt.kernel_options
    nps = np.asarray(desc.tensors_structure.number_of_points_per_surface, int)
    sp, op, og = ii.surface_points.sp_coords, ii.orientations.dip_positions, ii.orientations.dip_gradients
    nug_s, nug_o = ii.surface_points.nugget_effect_scalar, ii.orientations.nugget_effect_grad
    rng = np.random.default_rng(3)
    xyz = rng.uniform(-0.3, 0.3, size=(40, 3))

    def fields(sp_, op_, og_, pts, nug_s_=nug_s, nug_o_=nug_o):
        st = orc.prepare_stack(sp_, nug_s_, nps, op_, og_, nug_o_)
        w = orc.solve(orc.assemble_covariance(st, ko), orc.rhs(st, ko))
        return orc.evaluate(st, ko, w, pts, gradient=True)

    Z0, G0 = fields(sp, op, og, xyz)
    t = np.array([0.11, -0.07, 0.05])
    Z1, G1 = fields(sp + t, op + t, og, xyz + t)
    np.testing.assert_allclose(G1, G0, rtol=0, atol=1e-7 * np.abs(G0).max())
    dz = Z1 - Z0
    assert np.ptp(dz) < 1e-7 * np.ptp(Z0)                    # a constant shift at most
    # permute the non-reference points inside every surface, and the orientations
    starts = np.concatenate([[0], np.cumsum(nps)[:-1]])
    perm = np.arange(sp.shape[0])
    for s0, k in zip(starts, nps):
        perm[s0 + 1:s0 + k] = s0 + 1 + rng.permutation(k - 1)
    po = rng.permutation(op.shape[0])
    Z2, G2 = fields(sp[perm], op[po], og[po], xyz, nug_s[perm], nug_o[po])
    np.testing.assert_allclose(Z2, Z0, rtol=0, atol=1e-8 * np.abs(Z0).max())
    np.testing.assert_allclose(G2, G0, rtol=0, atol=1e-8 * np.abs(G0).max())


@pytest.mark.parametrize("kind", ["cubic", "exponential", "matern_5_2"])
def test_kernel_terms_are_self_consistent(kind):
    """The three covariance terms the assembly and the evaluation use, C, C'/r and C'', agree with numerical derivatives of
    C for every kernel (the exponential and Matern forms have no reference fixture: this at least ties their derivative
    terms to their own C), C(0) = 1, and the cubic kernel and its first two derivatives vanish at the range."""
    a = 1.7
    r = np.linspace(0.05, 1.6, 200)
    h = 1e-5
    C, Cp_r, Cpp = orc.kernel_terms(r, a, kind)
    Cp, _, _ = orc.kernel_terms(r + h, a, kind)
    Cm, _, _ = orc.kernel_terms(r - h, a, kind)
    np.testing.assert_allclose(Cp_r * r, (Cp - Cm) / (2 * h), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(Cpp, (Cp - 2 * C + Cm) / h ** 2, rtol=1e-4, atol=1e-5)
    assert abs(float(orc.kernel_terms(np.array([0.0]), a, kind)[0][0]) - 1.0) < 1e-15
    if kind == "cubic":
        Ca, Cpa, Cppa = orc.kernel_terms(np.array([a]), a, kind)
        assert abs(Ca[0]) < 1e-14 and abs(Cpa[0]) < 1e-14 and abs(Cppa[0]) < 1e-13


def test_matern_52_is_the_published_matern():
    """Matern covariance of smoothness nu = 5/2 in its general (Bessel) form, 2^(1-nu)/Gamma(nu) x^nu K_nu(x) with
    x = sqrt(2 nu) r / a, against the closed form the oracle (and the CUDA kernels) use."""
    from scipy.special import gamma, kv
    a, nu = 1.7, 2.5
    r = np.linspace(1e-3, 6.0, 400)
    x = np.sqrt(2 * nu) * r / a
    general = 2.0 ** (1 - nu) / gamma(nu) * x ** nu * kv(nu, x)
    np.testing.assert_allclose(orc.kernel_terms(r, a, "matern_5_2")[0], general, rtol=1e-12, atol=1e-15)
