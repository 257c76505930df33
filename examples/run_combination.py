#!/usr/bin/env python
"""Smallest end-to-end use of the backend (needs a B200): the reference's COMBINATION example model
(gempy/API/examples_generator.py:244-293) through the drop-in ``compute_model``.

    python examples/run_combination.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import examples as ex                      # noqa: E402
from gempy_b200.engine.compute import compute_model        # noqa: E402

model = ex.combination(refinement=6)
sol = compute_model(model.interpolation_input, model.options, model.descriptor)

print("octree levels (voxels):", [lvl.grid_centers.octree_grid.n_points for lvl in sol.octrees_output])
print("scalar field at the interfaces:", np.round(sol.scalar_field_at_surface_points, 6))
print("element order per group:", [o.tolist() for o in sol._ordered_elements])
lith = sol.raw_arrays.lith_block                              # finest-level regular lattice, filled from the octree
print("lith_block:", lith.shape, dict(zip(*np.unique(lith, return_counts=True))))
for name, mesh in zip(model.element_names, sol.dc_meshes):
    world = model.transform.apply_inverse(mesh.vertices)     # what GeoModel.solutions does (geo_model.py:117-118)
    print(f"mesh {name}: {mesh.vertices.shape[0]} vertices, {mesh.edges.shape[0]} triangles, z in "
          f"[{world[:, 2].min():.0f}, {world[:, 2].max():.0f}] m")
