"""CPU-only checks: the C-ABI library loads and exports every symbol include/gempy_b200.h declares, the ctypes
signatures cover the header, host-side logic (data model, example builders) behaves,
and nothing in the product package routes through the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "gempy_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpb_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from gempy_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/gempy_b200.h but not exported"
    # the ctypes table binds exactly the header's entry points
    assert sorted(_lib.SIGNATURES) == syms
    lib = _lib.lib()
    assert lib.gpb_version() == 200
    assert lib.gpb_launch_count() == 0          # no compute without a GPU


def test_struct_layouts_match_the_header():
    from gempy_b200 import _lib
    # gpb_stack: 6 ints, 4 doubles, 10 pointers; gpb_regular_grid: 6 doubles, 3 ints (+pad)
    assert ctypes.sizeof(_lib.GpbStack) == 6 * 4 + 4 * 8 + 10 * 8
    assert ctypes.sizeof(_lib.GpbRegularGrid) == 6 * 8 + 4 * 4


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gempy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)


def test_no_cpu_fallback_without_cuda():
    import torch
    from gempy_b200 import _lib, examples as ex
    from gempy_b200.engine import compute as gc
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.GpbError):
        gc.compute_model(*ex.horizontal_strat().args())


def test_example_models_match_the_reference_descriptors():
    from gempy_b200 import examples as ex
    from gempy_b200.engine.data import StackRelationType as R
    m = ex.combination()
    ii, opt, desc = m.args()
    # 102 SP (fault 3, rock3 15, rock2 39 + rock1 45), 8 ORI, base [4,2,2]
    assert ii.surface_points.n_points == 102 and ii.orientations.n_items == 8
    assert desc.stack_structure.number_of_points_per_stack.tolist() == [3, 15, 84]
    assert desc.stack_structure.number_of_orientations_per_stack.tolist() == [1, 1, 6]
    assert desc.tensors_structure.number_of_points_per_surface.tolist() == [3, 15, 39, 45]
    assert ii.grid.octree_grid.regular_grid_shape.tolist() == [4, 2, 2]
    assert [r for r in desc.stack_structure.masking_descriptor] == [R.FAULT, R.ERODE, R.BASEMENT]
    assert desc.stack_structure.faults_relations.astype(int).tolist() == [[0, 1, 1], [0, 0, 0], [0, 0, 0]]
    assert ii.unit_values.tolist() == [1, 2, 3, 4, 5]
    assert opt.number_octree_levels == 4 and opt.number_octree_levels_surface == 4
    h = ex.horizontal_strat()
    assert h.interpolation_input.grid.dense_grid.regular_grid_shape.tolist() == [50, 5, 50]
    assert h.interpolation_input.grid.dense_grid.n_points == 12500


def test_regular_grid_ordering_and_slices():
    from gempy_b200.engine.data import EngineGrid, GenericGrid, RegularGrid
    g = RegularGrid([0, 2, 0, 3, 0, 4], [2, 3, 4])
    v = g.values
    assert v.shape == (24, 3)
    np.testing.assert_allclose(v[0], [0.5, 0.5, 0.5])
    np.testing.assert_allclose(v[1], [0.5, 0.5, 1.5])       # z fastest
    np.testing.assert_allclose(v[4], [0.5, 1.5, 0.5])
    np.testing.assert_allclose(v[12], [1.5, 0.5, 0.5])      # x slowest
    eg = EngineGrid(octree_grid=RegularGrid([0, 1, 0, 1, 0, 1], [2, 2, 2]), dense_grid=g,
                    custom_grid=GenericGrid(np.zeros((5, 3))))
    assert eg.len_all_grids == 8 + 24 + 5
    assert eg.dense_grid_slice == slice(8, 32) and eg.custom_grid_slice == slice(32, 37)


def test_octree_to_regular_fill_rule():
    """The octree -> regular fill rule the device kernels (gpb_upsample2 + gpb_scatter_lattice) are checked against on the
    GPU (tests/test_gpu_parity.py): refined voxels are overwritten by their children, the others keep the parent value."""
    from oracle import gempy_oracle as orc
    base = np.array([2, 2, 2])
    lvl0 = {"lith": np.arange(8, dtype=float), "selected": np.array([1, 0, 0, 0, 0, 0, 0, 1], bool)}
    lvl1 = {"lith": 100 + np.arange(16, dtype=float), "selected": None}
    dense = orc.fill_regular_from_octree([lvl0, lvl1], base, lambda h: h["lith"]).reshape(4, 4, 4)
    # voxel 0 (i=j=k=0) was refined: its 8 children carry 100..107 in (x slow, z fast) order
    np.testing.assert_array_equal(dense[:2, :2, :2].ravel(), 100 + np.arange(8))
    np.testing.assert_array_equal(dense[2:, 2:, 2:].ravel(), 108 + np.arange(8))
    # an unrefined voxel keeps its parent value in all 8 cells
    assert (dense[:2, :2, 2:] == 1).all()


def test_interpolation_options_defaults_follow_the_serialization_golden():
    from gempy_b200.engine.data import InterpolationOptions
    o = InterpolationOptions.init_octree_options(refinement=3)
    k, e = o.kernel_options, o.evaluation_options
    assert (k.range, k.c_o, k.uni_degree, k.i_res, k.gi_res, k.number_dimensions) == (1.7, 10.0, 1, 4.0, 2.0, 3)
    assert k.kernel_function.name == "cubic" and k.kernel_solver == 1
    assert (e._number_octree_levels, e._number_octree_levels_surface, e.octree_curvature_threshold,
            e.octree_error_threshold, e.octree_min_level, e.evaluation_chunk_size) == (3, 4, -1.0, 1.0, 2, 500_000)
    assert o.sigmoid_slope == 5_000_000 and o.cache_mode == 3 and e.mesh_extraction is True
    assert o.number_octree_levels_surface == 3      # capped by the number of levels


def test_engine_inputs_npz_round_trip(tmp_path):
    """gempy_b200.engine.io: the three compute_model arguments survive the .npz form the bridge fixtures use."""
    from gempy_b200 import examples as ex
    from gempy_b200.engine.io import engine_inputs_from_npz, engine_inputs_to_npz
    m = ex.combination(refinement=3)
    p = str(tmp_path / "m.npz")
    engine_inputs_to_npz(p, *m.args())
    ii, opt, desc = engine_inputs_from_npz(p)
    np.testing.assert_array_equal(ii.surface_points.sp_coords, m.interpolation_input.surface_points.sp_coords)
    np.testing.assert_array_equal(ii.orientations.dip_gradients, m.interpolation_input.orientations.dip_gradients)
    np.testing.assert_array_equal(ii.grid.octree_grid.regular_grid_shape, [4, 2, 2])
    assert opt.number_octree_levels == 3 and opt.number_octree_levels_surface == 3
    assert [r.name for r in desc.stack_structure.masking_descriptor] == ["FAULT", "ERODE", "BASEMENT"]
    np.testing.assert_array_equal(desc.stack_structure.faults_relations, m.descriptor.stack_structure.faults_relations)


def test_shard_decision_cost_model(monkeypatch):
    """Per-level sharding policy (engine/compute.py:_shard_pays): BASELINE configs 3 and 5 shard their big levels over the
    ranks, the 15-stack multi-fault model and every shallow level are evaluated whole on each rank."""
    from gempy_b200.engine.compute import _shard_pays
    monkeypatch.delenv("GPB_SHARD_MIN_PAIRS", raising=False)
    assert _shard_pays(134_217_728, 5000, 1, 8) and _shard_pays(134_217_728, 5000, 1, 2)        # config 3, dense 512^3
    assert _shard_pays(10_500_000, 25_000, 1, 8)                                                # config 5, level 10
    assert not _shard_pays(6_000_000, 1870, 15, 8)                                              # config 4, level 8
    assert not _shard_pays(8 * 9 + 100, 25_000, 1, 8)                                           # any root level
    monkeypatch.setenv("GPB_SHARD_MIN_PAIRS", "0")
    assert _shard_pays(1, 1, 1, 2)
    monkeypatch.setenv("GPB_SHARD_MIN_PAIRS", "1e30")
    assert not _shard_pays(134_217_728, 5000, 1, 8)
