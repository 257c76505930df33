#!/usr/bin/env python
"""Field-only evaluation (no gradient) of the config-3 tables on a 256^3 regular grid (z-run kernel) and through the
config-4 compute_model (octet + point-list kernels): kernel time by CUDA events."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gempy_b200 import examples as ex                    # noqa: E402
from gempy_b200.engine import compute as gc              # noqa: E402

from gempy_b200.engine.data import AvailableKernelFunctions as K   # noqa: E402
kernel = sys.argv[1] if len(sys.argv) > 1 else "cubic"
eng = gc.B200Engine(0)
m = ex.synthetic_stress(n_sp_per_surface=1000, n_surfaces=4, n_ori=1000, resolution=(256, 256, 256), kernel=K[kernel])
ii, opt, desc = m.args()
with eng.hold_stream():
    st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device)
    w, _ = eng.solve_stack(st, "s")
    src = eng.pack(st, w)
    g = ii.grid.dense_grid
    seg = gc.Segment("dense", g.n_points, grid=gc.regular_descriptor(g), i0=0)
    Z = eng.empty(g.n_points)
    ts = []
    for r in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.evaluate_segment(st, src, seg, 0, Z, None, None)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    G = eng.empty(3, g.n_points)
    tg = []
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.evaluate_segment(st, src, seg, 0, Z, G, None)
        e1.record(); torch.cuda.synchronize(); tg.append(e0.elapsed_time(e1))
    del G
print(f"{kernel}: zrun field+gradient 256^3: min {min(tg[1:]):.3f} ms")
print(f"{kernel}: zrun field-only 256^3: min {min(ts[1:]):.3f} ms  median {np.median(ts[1:]):.3f} ms  checksum {float(Z.sum()):.12e}", flush=True)
del Z
mm = ex.synthetic_multi_fault(refinement=8) if kernel == 'cubic' else None
if mm is not None:
    ws = []
    for r in range(8):
        mm = ex.synthetic_multi_fault(refinement=8)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sol = gc.compute_model(*mm.args(), engine=eng)
        torch.cuda.synchronize(); ws.append(time.perf_counter() - t0)
        del sol
    print(f"config-4 compute_model: min {min(ws[2:]) * 1e3:.2f} ms  median {np.median(ws[2:]) * 1e3:.2f} ms", flush=True)
