#!/usr/bin/env python
"""Where the compute_model wall goes on the multi-fault octree model (BASELINE configs[3] shape): host wall, device time
between two events on the stream, launches, and the host reads one by one."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import _lib, examples as ex           # noqa: E402
from gempy_b200.engine import compute as gc           # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=8)
ap.add_argument("--model", default="multi_fault")
args = ap.parse_args()
eng = gc.B200Engine(0)
build = (lambda: ex.synthetic_multi_fault(refinement=args.levels)) if args.model == "multi_fault" else (lambda: ex.combination(refinement=args.levels))
gc.compute_model(*build().args(), engine=eng)
rec = {"model": args.model, "levels": args.levels}
walls, devs = [], []
for _ in range(5):
    m = build()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    e0.record()
    sol = gc.compute_model(*m.args(), engine=eng)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    walls.append(t2 - t0)
    devs.append(e0.elapsed_time(e1) * 1e-3)
    rec["host_enqueue_s"] = t1 - t0
    rec["launches"] = _lib.launch_count() - l0
rec["wall_s_min"] = min(walls)
rec["device_span_s_min"] = min(devs)
t = time.perf_counter(); lb = sol.raw_arrays.lith_block; rec["read_lith_block_s"] = time.perf_counter() - t
t = time.perf_counter(); fb = sol.raw_arrays.fault_block; rec["read_fault_block_s"] = time.perf_counter() - t
if sol.dc_meshes:
    t = time.perf_counter(); nv = sum(mm.vertices.shape[0] for mm in sol.dc_meshes); rec["read_vertices_s"] = time.perf_counter() - t
    t = time.perf_counter(); nt = sum(mm.edges.shape[0] for mm in sol.dc_meshes); rec["read_edges_s"] = time.perf_counter() - t
    rec["n_vertices"], rec["n_triangles"] = int(nv), int(nt)
t = time.perf_counter(); z = sol.octrees_output[-1].outputs_centers[-1].exported_fields.scalar_field; rec["read_last_level_field_s"] = time.perf_counter() - t
rec["last_level_points"] = int(z.shape[0])
print(json.dumps(rec))
