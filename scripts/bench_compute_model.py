#!/usr/bin/env python
"""Metric M2 (BASELINE.json): compute_model wall seconds, B200 backend vs the restated reference (numpy oracle) on
the same model, plus parity of the two results.  Configs: BASELINE configs[1] (COMBINATION, octree level 6) and a
multi-fault synthetic model (configs[3] shape at a CPU-feasible octree depth)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import examples as ex                 # noqa: E402
from gempy_b200.engine import compute as gc           # noqa: E402
from oracle import gempy_oracle as orc                # noqa: E402


def run(name, build, eng, cpu=True, reps=3):
    sol = gc.compute_model(*build().args(), engine=eng)          # warm-up (allocator, module load)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol = gc.compute_model(*build().args(), engine=eng)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    rec = {"model": name, "b200_wall_s": min(ts), "levels": [int(l.grid_centers.octree_grid.values.shape[0]) for l in sol.octrees_output],
           "n_meshes": None if sol.dc_meshes is None else len(sol.dc_meshes),
           "n_vertices": None if sol.dc_meshes is None else int(sum(m.vertices.shape[0] for m in sol.dc_meshes))}
    if cpu:
        t0 = time.perf_counter()
        ref = orc.compute_model(*build().args())
        rec["oracle_numpy_wall_s"] = time.perf_counter() - t0
        rec["speedup"] = rec["oracle_numpy_wall_s"] / rec["b200_wall_s"]
        worst, ids_ok, n_ids = 0.0, 0, 0
        for a, b in zip(sol.octrees_output, ref.levels):
            nv = b.centers.shape[0]
            assert a.grid_centers.octree_grid.values.shape[0] == nv, "leaf lists differ"
            for oa, ob in zip(a.outputs_centers, b.fields.stacks):
                za, zb = oa.exported_fields.scalar_field[:nv], ob.Z[:nv]
                worst = max(worst, float(np.abs(za - zb).max() / max(np.abs(zb).max(), 1e-300)))
            near = np.zeros(nv, bool)
            for ob in b.fields.stacks:
                near |= (np.abs(ob.Z[:nv, None] - ob.isovalues[None, :]) < 1e-6).any(axis=1)
            ia, ib = np.rint(a.outputs_centers[-1].block[:nv]), b.fields.lith_ids[:nv]
            ids_ok += int((ia[~near] == ib[~near]).sum())
            n_ids += int((~near).sum())
        rec["max_rel_field_diff"] = worst
        rec["lith_ids_exact"] = f"{ids_ok}/{n_ids}"
        if ref.meshes:
            dv = max(float(np.abs(a.vertices - b.vertices).max()) for a, b in zip(sol.dc_meshes, ref.meshes) if a.vertices.shape == b.vertices.shape)
            rec["max_vertex_diff"] = dv
            rec["mesh_shapes_equal"] = all(a.vertices.shape == b.vertices.shape for a, b in zip(sol.dc_meshes, ref.meshes))
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--deep", action="store_true", help="also run the GPU-only deep configurations")
    args = ap.parse_args()
    eng = gc.B200Engine(0)
    cpu = not args.no_cpu
    run("combination_octree6 (BASELINE configs[1])", lambda: ex.combination(refinement=6), eng, cpu)
    run("multi_fault_10f_5s_octree4 (BASELINE configs[3] shape, CPU-feasible depth)",
        lambda: ex.synthetic_multi_fault(refinement=4), eng, cpu)
    if args.deep:
        run("multi_fault_10f_5s_octree6", lambda: ex.synthetic_multi_fault(refinement=6), eng, False, reps=1)
        run("multi_fault_10f_5s_octree8 (BASELINE configs[3])", lambda: ex.synthetic_multi_fault(refinement=8), eng, False, reps=1)


if __name__ == "__main__":
    main()
