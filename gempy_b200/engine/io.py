"""Engine inputs <-> .npz: the three arguments of ``compute_model`` (interpolation input, options, data descriptor) as plain
arrays.  Used to carry the inputs the reference's bridge builds (gempy/modules/data_manipulation/_engine_factory.py:14-58)
to a machine that has the GPU but not the reference tree (tests/golden/bridge_*.npz)."""
from __future__ import annotations

import json
from typing import Tuple

import numpy as np

from .data import (AvailableKernelFunctions, BlockSolutionType, EngineGrid, GenericGrid, InputDataDescriptor,
                   InterpolationInput, InterpolationOptions, Orientations, RegularGrid, StackRelationType, StacksStructure,
                   SurfacePoints, TensorsStructure)


def _rel_code(r) -> int:
    if r is False or r is None:
        return StackRelationType.BASEMENT.value
    return int(getattr(r, "value", r))


def engine_inputs_to_npz(path: str, ii: InterpolationInput, options: InterpolationOptions, desc: InputDataDescriptor) -> None:
    ko, eo = options.kernel_options, options.evaluation_options
    ss, ts = desc.stack_structure, desc.tensors_structure
    opt = {"range": float(ko.range), "c_o": float(ko.c_o), "uni_degree": int(ko.uni_degree), "i_res": float(ko.i_res),
           "gi_res": float(ko.gi_res), "kernel_function": getattr(ko.kernel_function, "name", str(ko.kernel_function)),
           "number_octree_levels": int(eo.number_octree_levels), "number_octree_levels_surface": int(eo._number_octree_levels_surface),
           "octree_min_level": int(eo.octree_min_level), "mesh_extraction": bool(eo.mesh_extraction),
           "compute_scalar_gradient": bool(eo.compute_scalar_gradient), "sigmoid_slope": float(options.sigmoid_slope),
           "block_solutions_type": getattr(options.block_solutions_type, "name", str(options.block_solutions_type))}
    g = ii.grid
    arrays = {
        "sp_coords": ii.surface_points.sp_coords, "sp_nugget": ii.surface_points.nugget_effect_scalar,
        "ori_pos": ii.orientations.dip_positions, "ori_grad": ii.orientations.dip_gradients,
        "ori_nugget": ii.orientations.nugget_effect_grad, "unit_values": np.asarray(ii.unit_values),
        "octree_extent": g.octree_grid.orthogonal_extent, "octree_shape": g.octree_grid.regular_grid_shape,
        "points_per_surface": ts.number_of_points_per_surface, "points_per_stack": ss.number_of_points_per_stack,
        "orientations_per_stack": ss.number_of_orientations_per_stack, "surfaces_per_stack": ss.number_of_surfaces_per_stack,
        "relations": np.array([_rel_code(r) for r in ss.masking_descriptor]),
        "faults_relations": np.zeros((0, 0), bool) if ss.faults_relations is None else np.asarray(ss.faults_relations, bool),
        "options_json": np.frombuffer(json.dumps(opt).encode(), dtype=np.uint8),
    }
    if g.dense_grid is not None:
        arrays["dense_extent"], arrays["dense_shape"] = g.dense_grid.orthogonal_extent, g.dense_grid.regular_grid_shape
    for name in ("custom_grid", "topography", "sections"):
        gg = getattr(g, name)
        if gg is not None:
            arrays[name] = gg.values
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})


def engine_inputs_from_npz(path: str) -> Tuple[InterpolationInput, InterpolationOptions, InputDataDescriptor]:
    z = np.load(path)
    opt = json.loads(bytes(z["options_json"]).decode())
    options = InterpolationOptions.from_args(range=opt["range"], c_o=opt["c_o"], uni_degree=opt["uni_degree"], i_res=opt["i_res"],
                                             gi_res=opt["gi_res"], number_octree_levels=opt["number_octree_levels"],
                                             kernel_function=AvailableKernelFunctions[opt["kernel_function"]],
                                             mesh_extraction=opt["mesh_extraction"],
                                             compute_scalar_gradient=opt["compute_scalar_gradient"], sigmoid_slope=opt["sigmoid_slope"])
    options.evaluation_options.number_octree_levels_surface = opt["number_octree_levels_surface"]
    options.evaluation_options.octree_min_level = opt["octree_min_level"]
    options.block_solutions_type = BlockSolutionType[opt["block_solutions_type"]]
    grid = EngineGrid(octree_grid=RegularGrid(z["octree_extent"], z["octree_shape"]),
                      dense_grid=RegularGrid(z["dense_extent"], z["dense_shape"]) if "dense_extent" in z else None,
                      custom_grid=GenericGrid(z["custom_grid"]) if "custom_grid" in z else None,
                      topography=GenericGrid(z["topography"]) if "topography" in z else None,
                      sections=GenericGrid(z["sections"]) if "sections" in z else None)
    ii = InterpolationInput(SurfacePoints(z["sp_coords"], z["sp_nugget"]), Orientations(z["ori_pos"], z["ori_grad"], z["ori_nugget"]),
                            grid, unit_values=z["unit_values"], weights=[])
    fr = z["faults_relations"]
    desc = InputDataDescriptor(TensorsStructure(z["points_per_surface"]),
                               StacksStructure(z["points_per_stack"], z["orientations_per_stack"], z["surfaces_per_stack"],
                                               [StackRelationType(int(r)) for r in z["relations"]],
                                               faults_relations=fr if fr.size else None))
    return ii, options, desc
