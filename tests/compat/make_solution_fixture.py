#!/usr/bin/env python
"""GPU box: compute COMBINATION from the reference bridge's inputs (tests/golden/bridge_combination.npz) and save the small
host arrays GeoModel.solutions reads (geo_model.py:100-127) -> gpurun_out/solution_combination.npz (copied to tests/golden/
by hand; consumed on the CPU side by tests/test_compat_reference.py::test_geomodel_solutions_setter_consumes_backend_solutions)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gempy_b200.engine import compute as gc                 # noqa: E402
from gempy_b200.engine.io import engine_inputs_from_npz     # noqa: E402

sol = gc.compute_model(*engine_inputs_from_npz(os.path.join(ROOT, "tests", "golden", "bridge_combination.npz")))
out = {"scalar_field_at_surface_points": np.asarray(sol.scalar_field_at_surface_points), "n_meshes": len(sol.dc_meshes),
       "n_groups": len(sol._ordered_elements)}
for k, m in enumerate(sol.dc_meshes):
    out[f"vertices_{k}"], out[f"edges_{k}"] = m.vertices, m.edges
for g, o in enumerate(sol._ordered_elements):
    out[f"order_{g}"] = np.asarray(o)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "solution_combination.npz"), **out)
print("saved", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
