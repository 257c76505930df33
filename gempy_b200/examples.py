"""Example and synthetic models, built the way GemPy's API builds the engine inputs.

Restates, for the models BASELINE.json's configs name, the host-side steps of
gempy/API/examples_generator.py:132-293 + gempy/API/initialization_API.py:21-104 +
gempy/modules/data_manipulation/_engine_factory.py:14-105:

  tables (CSV rows, alphabetical element ids: gempy/core/data/_data_points_helpers.py:15-16)
  -> stacks via ``map_stack_to_surfaces`` (gempy/API/map_stack_to_surfaces_API.py:10-75)
  -> input transform ``Transform.from_input_points`` (gempy/core/data/geo_model.py:244-247)
  -> azimuth/dip/polarity -> gradient (gempy/API/io_API.py:96-103)
  -> octree base resolution (gempy/core/data/grid.py:127-151)
  -> InterpolationInput / InterpolationOptions / InputDataDescriptor.

The input tables are data shipped in ``gempy_b200/data/example_inputs.json`` (extracted from the
reference's example CSVs by tests/golden/make_fixtures.py).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .engine.data import (CenteredGrid, EngineGrid, GenericGrid, InputDataDescriptor, InterpolationInput, InterpolationOptions,
                          Orientations, RegularGrid, StackRelationType, StacksStructure, SurfacePoints,
                          TensorsStructure, Transform, BlockSolutionType, AvailableKernelFunctions)

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "example_inputs.json")

DEFAULT_SP_NUGGET = 2e-5      # gempy/core/data/surface_points.py:11
DEFAULT_ORI_NUGGET = 0.01     # gempy/core/data/orientations.py:11


@dataclass
class ExampleModel:
    """Everything ``compute_model`` needs plus what a test needs to go back to world coordinates."""
    name: str
    interpolation_input: InterpolationInput
    options: InterpolationOptions
    descriptor: InputDataDescriptor
    transform: Transform
    extent: np.ndarray
    element_names: List[str]

    def args(self):
        return self.interpolation_input, self.options, self.descriptor


def _tables(model: str):
    with open(_DATA) as fh:
        return json.load(fh)[model]


def _gradients(az, dip, pol):
    az, dip, pol = (np.asarray(v, float) for v in (az, dip, pol))
    gx = np.sin(np.deg2rad(dip)) * np.sin(np.deg2rad(az)) * pol
    gy = np.sin(np.deg2rad(dip)) * np.cos(np.deg2rad(az)) * pol
    gz = np.cos(np.deg2rad(dip)) * pol
    return np.stack([gx, gy, gz], axis=1)


def octree_base_resolution(extent: Sequence[float], legacy: bool = False) -> np.ndarray:
    """gempy/core/data/grid.py:133-142 (np.round is banker's rounding: 2.5 -> 2)."""
    if legacy:
        return np.array([2, 2, 2])
    e = np.asarray(extent, float)
    lengths = np.array([e[1] - e[0], e[3] - e[2], e[5] - e[4]])
    return np.round(lengths / lengths.min()).astype(int) * 2


def build_model(name: str, sp_xyz: Dict[str, np.ndarray], ori_xyz: Dict[str, np.ndarray],
                ori_grad: Dict[str, np.ndarray], stacks: Sequence[Tuple[str, Sequence[str], StackRelationType]],
                extent: Sequence[float], *, refinement: Optional[int] = None,
                resolution: Optional[Sequence[int]] = None, fault_relations: Optional[np.ndarray] = None,
                custom_xyz: Optional[np.ndarray] = None, options: Optional[InterpolationOptions] = None,
                transform: Optional[Transform] = None, sp_nugget: float = DEFAULT_SP_NUGGET,
                ori_nugget: float = DEFAULT_ORI_NUGGET, legacy_octree_init: bool = False,
                centered: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None) -> ExampleModel:
    """``stacks`` = [(group name, [element names in order], relation)] from youngest to oldest."""
    elements = [e for _, els, _ in stacks for e in els]
    sp = np.concatenate([np.asarray(sp_xyz[e], float).reshape(-1, 3) for e in elements])
    empty = np.zeros((0, 3))
    op = np.concatenate([np.asarray(ori_xyz.get(e, empty), float).reshape(-1, 3) for e in elements])
    og = np.concatenate([np.asarray(ori_grad.get(e, empty), float).reshape(-1, 3) for e in elements])
    extent = np.asarray(extent, float)
    if transform is None:
        transform = Transform.from_input_points(sp, op)
    sp_t, op_t, og_t = transform.apply(sp), transform.apply(op), transform.transform_gradient(og)
    corners = np.array([[extent[i], extent[2 + j], extent[4 + k]] for i in (0, 1) for j in (0, 1) for k in (0, 1)])
    ct = transform.apply(corners)
    ext_t = np.array([ct[:, 0].min(), ct[:, 0].max(), ct[:, 1].min(), ct[:, 1].max(), ct[:, 2].min(), ct[:, 2].max()])

    dense = None
    if resolution is not None:
        dense = RegularGrid(ext_t, np.asarray(resolution, int))
        base = np.array([2, 2, 2])
        if options is None:
            options = InterpolationOptions.init_dense_grid_options()
    else:
        base = octree_base_resolution(extent, legacy_octree_init)
        if options is None:
            options = InterpolationOptions.init_octree_options(refinement=refinement or 1)
    octree = RegularGrid(ext_t, base)
    custom = GenericGrid(transform.apply(custom_xyz)) if custom_xyz is not None else None
    geo = None
    if centered is not None:                      # _engine_factory.py:82-87
        c_xyz, c_radius, c_res = centered
        geo = CenteredGrid(transform.apply(np.asarray(c_xyz, float).reshape(-1, 3)),
                           transform.scale_points(np.atleast_2d(np.asarray(c_radius, float)))[0], np.asarray(c_res, int))
    grid = EngineGrid(octree_grid=octree, dense_grid=dense, custom_grid=custom, geophysics_grid=geo)
    options.block_solutions_type = BlockSolutionType.DENSE_GRID if dense is not None else BlockSolutionType.OCTREE

    n_elem = len(elements)
    ii = InterpolationInput(
        surface_points=SurfacePoints(sp_t, sp_nugget),
        orientations=Orientations(op_t, og_t, ori_nugget),
        grid=grid,
        unit_values=np.arange(n_elem + 1) + 1,      # + basement (structural_frame.py:367-370)
        weights=[],
    )
    rel = [r for _, _, r in stacks]
    rel[-1] = StackRelationType.BASEMENT             # structural_frame.py:325-330
    n_st = len(stacks)
    fr = np.zeros((n_st, n_st), bool) if fault_relations is None else np.asarray(fault_relations, bool)
    desc = InputDataDescriptor(
        TensorsStructure(np.array([np.asarray(sp_xyz[e]).reshape(-1, 3).shape[0] for e in elements])),
        StacksStructure(
            number_of_points_per_stack=np.array([sum(np.asarray(sp_xyz[e]).reshape(-1, 3).shape[0] for e in els) for _, els, _ in stacks]),
            number_of_orientations_per_stack=np.array([sum(np.asarray(ori_xyz.get(e, empty)).reshape(-1, 3).shape[0] for e in els) for _, els, _ in stacks]),
            number_of_surfaces_per_stack=np.array([len(els) for _, els, _ in stacks]),
            masking_descriptor=rel,
            faults_relations=fr,
        ))
    return ExampleModel(name, ii, options, desc, transform, extent, elements)


def _from_tables(model: str):
    t = _tables(model)
    sp, op, og = {}, {}, {}
    f = np.array(t["surface_points"]["formation"])
    xyz = np.stack([t["surface_points"][k] for k in "XYZ"], axis=1).astype(float)
    for name in np.unique(f):
        sp[str(name)] = xyz[f == name]
    o = t["orientations"]
    fo = np.array(o["formation"])
    oxyz = np.stack([o[k] for k in "XYZ"], axis=1).astype(float)
    g = _gradients(o["azimuth"], o["dip"], o["polarity"])
    for name in np.unique(fo):
        op[str(name)] = oxyz[fo == name]
        og[str(name)] = g[fo == name]
    return sp, op, og


E, F = StackRelationType.ERODE, StackRelationType.FAULT


def horizontal_strat(resolution=(50, 5, 50), **kw) -> ExampleModel:
    """BASELINE config 1 (examples_generator.py:132-162): model1, dense 50x5x50 AND refinement=3 -- create_geomodel
    (initialization_API.py:67-85) builds the dense grid but keeps the default octree options, so the engine runs three
    octree levels on the [2,2,2] root next to the dense grid and extracts meshes; the block solution type is inferred
    DENSE_GRID (geo_model.py:324-339).  Checked against the reference's own bridge in tests/test_compat_reference.py."""
    sp, op, og = _from_tables("model1")
    kw.setdefault("options", InterpolationOptions.init_octree_options(refinement=3))
    m = build_model("horizontal", sp, op, og, [("Strat_Series", ["rock2", "rock1"], E)],
                    [0, 1000, 0, 1000, 0, 1000], resolution=resolution, **kw)
    return m


def anticline(refinement: int = 5, **kw) -> ExampleModel:
    """examples_generator.py:165-194: model2, octree refinement 5."""
    sp, op, og = _from_tables("model2")
    return build_model("fold", sp, op, og, [("Strat_Series", ["rock2", "rock1"], E)],
                       [0, 1000, 0, 1000, 0, 1000], refinement=refinement, **kw)


def one_fault(refinement: int = 6, **kw) -> ExampleModel:
    """examples_generator.py:197-241: model5, fault stack offsets the strat stack."""
    sp, op, og = _from_tables("model5")
    return build_model("fault", sp, op, og,
                       [("Fault_Series", ["fault"], F), ("Strat_Series", ["rock2", "rock1"], E)],
                       [0, 1000, 0, 1000, 0, 1000], refinement=refinement,
                       fault_relations=np.array([[0, 1], [0, 0]]), **kw)


def combination(refinement: int = 4, **kw) -> ExampleModel:
    """BASELINE config 2 (examples_generator.py:244-293): model7, fault + unconformity + fold,
    octree base [4,2,2]."""
    sp, op, og = _from_tables("model7")
    m = build_model("combination", sp, op, og,
                    [("Fault_Series", ["fault"], F), ("Strat_Series1", ["rock3"], E),
                     ("Strat_Series2", ["rock2", "rock1"], E)],
                    [0, 2500, 0, 1000, 0, 1000], refinement=refinement,
                    fault_relations=np.array([[0, 1, 1], [0, 0, 0], [0, 0, 0]]), **kw)
    m.options.evaluation_options.number_octree_levels_surface = 4
    return m


def two_layers_gravity(resolution=(500, 1, 500)):
    """The model of test/test_modules/test_geophysics/test_gravity.py:10-89: two horizontal surfaces, one device at
    (6, 0, 4) with a [10, 10, 100] kernel of radius 16000, densities [2.6, 2.4, 3.2].  Returns (model, geophysics_input);
    the reference's known answer is gravity = [-1624.1714]."""
    from .engine.geophysics import GeophysicsInput, calculate_gravity_gradient
    sp = {"surface1": np.array([[3, 0, 3.05], [9, 0, 3.05]]), "surface2": np.array([[3, 0, 1.02], [9, 0, 1.02]])}
    op = {"surface1": np.array([[6.0, 0.0, 4.0]])}
    og = {"surface1": np.array([[0.0, 0.0, 1.0]])}
    centers, radius, res = np.array([[6.0, 0.0, 4.0]]), np.array([16000.0, 16000.0, 16000.0]), np.array([10, 10, 100])
    m = build_model("2-layers", sp, op, og, [("default", ["surface1", "surface2"], E)], [0, 12, -2, 2, 0, 4],
                    resolution=resolution, centered=(centers, radius, res))
    tz = calculate_gravity_gradient(CenteredGrid(centers, radius, res))     # real coordinates, as gp.calculate_gravity_gradient
    return m, GeophysicsInput(tz=tz, densities=np.array([2.6, 2.4, 3.2]))


def greenstone(refinement: Optional[int] = None, **kw) -> ExampleModel:
    """The Greenstone model the reference ships as examples/data/gempy_models/Greenstone.gempy
    (gempy/API/examples_generator.py:489-508): 3 series (EarlyGranite | SimpleMafic2, SimpleBIF | SimpleMafic1), 70 surface
    points, 41 orientations, extent 51 x 67 x 20 km, octree 6 levels; the stored input transform is used as is."""
    with open(os.path.join(os.path.dirname(_DATA), "greenstone.json")) as fh:
        d = json.load(fh)
    sp, op, og, stacks = {}, {}, {}, []
    for g in d["groups"]:
        names = []
        for e in g["elements"]:
            nm = e["name"]
            sp[nm] = np.asarray(e["sp_xyz"], float).reshape(-1, 3)
            op[nm] = np.asarray(e["ori_xyz"], float).reshape(-1, 3)
            og[nm] = np.asarray(e["ori_grad"], float).reshape(-1, 3)
            names.append(nm)
        stacks.append((g["name"], names, StackRelationType(g["structural_relation"])))
    t = d["input_transform"]
    tr = Transform(np.asarray(t["position"], float), np.asarray(t["rotation"], float), np.asarray(t["scale"], float))
    return build_model("Greenstone", sp, op, og, stacks, d["extent"], refinement=refinement or d["number_octree_levels"],
                       transform=tr, legacy_octree_init=True, **kw)


# ------------------------------------------------------------------------------- synthetic configs
def synthetic_stress(n_sp_per_surface: int = 1000, n_surfaces: int = 4, n_ori: int = 1000,
                     resolution: Sequence[int] = (512, 512, 512), seed: int = 1234,
                     kernel: AvailableKernelFunctions = AvailableKernelFunctions.cubic,
                     refinement: Optional[int] = None) -> ExampleModel:
    """BASELINE config 3 / 5: single stack, surfaces
    ``z_k = 0.15 (k - (n-1)/2) + 0.05 sin(2 pi x) cos(2 pi y)`` sampled at uniform (x, y) in [-0.4, 0.4]^2 in
    *transformed* space, orientations = analytic unit normals at uniform points; grid over [-0.5, 0.5]^3.
    The data is generated directly in transformed coordinates (identity transform)."""
    rng = np.random.default_rng(seed)
    sp, op, og = {}, {}, {}
    names = [f"surface{k}" for k in range(n_surfaces)]

    def height(k, x, y):
        return 0.15 * ((n_surfaces - 1) / 2 - k) + 0.05 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)

    per_ori = [n_ori // n_surfaces + (1 if k < n_ori % n_surfaces else 0) for k in range(n_surfaces)]
    for k, nm in enumerate(names):
        xy = rng.uniform(-0.4, 0.4, size=(n_sp_per_surface, 2))
        sp[nm] = np.column_stack([xy, height(k, xy[:, 0], xy[:, 1])])
        xo = rng.uniform(-0.4, 0.4, size=(per_ori[k], 2))
        dzdx = 0.05 * 2 * np.pi * np.cos(2 * np.pi * xo[:, 0]) * np.cos(2 * np.pi * xo[:, 1])
        dzdy = -0.05 * 2 * np.pi * np.sin(2 * np.pi * xo[:, 0]) * np.sin(2 * np.pi * xo[:, 1])
        nrm = np.column_stack([-dzdx, -dzdy, np.ones_like(dzdx)])
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        op[nm] = np.column_stack([xo, height(k, xo[:, 0], xo[:, 1])])
        og[nm] = nrm
    ident = Transform(np.zeros(3), np.zeros(3), np.ones(3))
    m = build_model("synthetic_stress", sp, op, og, [("Stack", names, E)], [-0.5, 0.5, -0.5, 0.5, -0.5, 0.5],
                    resolution=None if refinement else resolution, refinement=refinement, transform=ident)
    m.options.kernel_options.kernel_function = kernel
    return m


def synthetic_multi_fault(n_faults: int = 10, n_series: int = 5, surfaces_per_series: int = 3,
                          n_sp_per_surface: int = 100, n_ori_per_series: int = 30, n_sp_fault: int = 20,
                          n_ori_fault: int = 2, refinement: int = 8, seed: int = 1234) -> ExampleModel:
    """BASELINE config 4: planar fault stacks with random strike/dip + stratigraphic series
    carrying one fault-drift column per fault."""
    rng = np.random.default_rng(seed)
    sp, op, og = {}, {}, {}
    strat_stacks = []
    zs = np.linspace(0.35, -0.35, n_series * surfaces_per_series)
    for s in range(n_series):
        names = []
        amp, ph = rng.uniform(0.02, 0.05), rng.uniform(0, 2 * np.pi)

        def height(z0, x, y):
            return z0 + amp * np.sin(2 * np.pi * x + ph) * np.cos(2 * np.pi * y)

        for k in range(surfaces_per_series):
            nm = f"s{s}_rock{k}"
            names.append(nm)
            z0 = zs[s * surfaces_per_series + k]
            xy = rng.uniform(-0.45, 0.45, size=(n_sp_per_surface, 2))
            sp[nm] = np.column_stack([xy, height(z0, xy[:, 0], xy[:, 1])])
            no = n_ori_per_series // surfaces_per_series
            xo = rng.uniform(-0.4, 0.4, size=(no, 2))
            dzdx = amp * 2 * np.pi * np.cos(2 * np.pi * xo[:, 0] + ph) * np.cos(2 * np.pi * xo[:, 1])
            dzdy = -amp * 2 * np.pi * np.sin(2 * np.pi * xo[:, 0] + ph) * np.sin(2 * np.pi * xo[:, 1])
            nrm = np.column_stack([-dzdx, -dzdy, np.ones_like(dzdx)])
            nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
            op[nm] = np.column_stack([xo, height(z0, xo[:, 0], xo[:, 1])])
            og[nm] = nrm
        strat_stacks.append((f"Series{s}", names, E))

    def increments(side):
        """Rows rest - ref of a 0/1 side function, per series: what a fault-drift column looks like."""
        cols = []
        for _, names, _ in strat_stacks:
            cols.append(np.concatenate([side[nm][1:] - side[nm][0] for nm in names]))
        return cols

    fault_stacks, accepted = [], [[] for _ in strat_stacks]
    f = 0
    while f < n_faults:
        strike = rng.uniform(0, np.pi)
        dip = rng.uniform(np.deg2rad(55), np.deg2rad(85))
        n = np.array([np.cos(strike) * np.sin(dip), np.sin(strike) * np.sin(dip), np.cos(dip)])
        c = rng.uniform(-0.3, 0.3, size=3) * np.array([1, 1, 0.2])
        # the fault must split every series' surface points in a way no earlier fault (or combination) already does,
        # otherwise two fault-drift columns coincide and the co-kriging system is singular
        side = {nm: ((sp[nm] - c) @ n > 0).astype(float) for _, names, _ in strat_stacks for nm in names}
        cols = increments(side)
        ok = True
        for k, col in enumerate(cols):
            M = np.stack(accepted[k] + [col], axis=1)
            if np.linalg.matrix_rank(M) < M.shape[1]:
                ok = False
        if not ok:
            continue
        for k, col in enumerate(cols):
            accepted[k].append(col)
        nm = f"fault{f}"
        u = np.cross(n, [0, 0, 1.0]); u /= np.linalg.norm(u)
        v = np.cross(n, u)
        ab = rng.uniform(-0.4, 0.4, size=(n_sp_fault, 2))
        sp[nm] = c + ab[:, :1] * u + ab[:, 1:] * v
        ab = rng.uniform(-0.3, 0.3, size=(n_ori_fault, 2))
        op[nm] = c + ab[:, :1] * u + ab[:, 1:] * v
        og[nm] = np.repeat(n[None], n_ori_fault, 0)
        fault_stacks.append((f"Fault{f}", [nm], F))
        f += 1
    stacks = fault_stacks + strat_stacks
    n_st = n_faults + n_series
    fr = np.zeros((n_st, n_st), bool)
    for f in range(n_faults):
        fr[f, n_faults:] = True       # every fault offsets every stratigraphic series (fault planes span the model,
                                      # so each series has surface points on both sides and the drift columns are
                                      # independent); faults do not offset each other
    ident = Transform(np.zeros(3), np.zeros(3), np.ones(3))
    return build_model("synthetic_multi_fault", sp, op, og, stacks, [-0.5, 0.5, -0.5, 0.5, -0.5, 0.5],
                       refinement=refinement, fault_relations=fr, transform=ident, legacy_octree_init=True)
