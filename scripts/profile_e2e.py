#!/usr/bin/env python
"""Where does the end-to-end step of bench.py go?  Stage walls of compute_dense_fields on config 3 (512^3)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import examples as ex
from gempy_b200.engine import compute as gc

eng = gc.B200Engine(0)
m = ex.synthetic_stress(n_sp_per_surface=1000, n_surfaces=4, n_ori=1000, resolution=(512, 512, 512))
ii, opt, desc = m.args()
npts = ii.grid.dense_grid.n_points
out = torch.empty((4, npts), dtype=torch.float64, pin_memory=True)
def sync(): torch.cuda.synchronize()
for n_slabs in (8, 16, 32):
    gc.compute_dense_fields(ii, opt, desc, engine=eng, out=out, n_slabs=n_slabs); sync()
    t0 = time.perf_counter()
    gc.compute_dense_fields(ii, opt, desc, engine=eng, out=out, n_slabs=n_slabs); sync()
    print(json.dumps({"n_slabs": n_slabs, "e2e_s": time.perf_counter() - t0}))
# stage by stage
rec = {}
sync(); t = time.perf_counter()
st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device); sync(); rec["tables_h2d_s"] = time.perf_counter() - t; t = time.perf_counter()
w, path = eng.solve_stack(st); sync(); rec["assemble_plus_solve_s"] = time.perf_counter() - t; rec["solver_path"] = path; t = time.perf_counter()
src = eng.pack(st, w); sync(); rec["pack_s"] = time.perf_counter() - t; t = time.perf_counter()
buf = eng.empty(4, npts // 8); sync(); rec["alloc_slab_s"] = time.perf_counter() - t; t = time.perf_counter()
seg = gc.Segment("dense_grid", npts // 8, grid=gc.regular_descriptor(ii.grid.dense_grid), i0=0)
eng.evaluate_segment(st, src, seg, 0, buf[0], buf[1:], None); sync(); rec["eval_one_slab_s"] = time.perf_counter() - t; t = time.perf_counter()
for a in range(4):
    out[a, :npts // 8].copy_(buf[a], non_blocking=True)
sync(); rec["d2h_one_slab_s"] = time.perf_counter() - t
rec["d2h_gbs"] = 4 * (npts // 8) * 8 / rec["d2h_one_slab_s"] / 1e9
print(json.dumps(rec))
# the solve alone, wall clock with a synchronize on both sides vs. host enqueue time
for rep in range(4):
    sync()
    t = time.perf_counter()
    w, _ = eng.solve_stack(st)
    t_enq = time.perf_counter() - t
    sync()
    print(json.dumps({"assemble_plus_solve_wall_s": time.perf_counter() - t, "host_side_s": t_enq}))
