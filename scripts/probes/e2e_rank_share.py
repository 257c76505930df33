#!/usr/bin/env python
"""One rank's share of the 512^3 benchmark step at N ranks, timed on ONE GPU: compute_dense_fields over the first 1/N of the
grid (host in, pinned host out), wave-aligned slabs against plane-aligned ones.  `python scripts/probes/e2e_rank_share.py 8`"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gempy_b200 import examples as ex                    # noqa: E402
from gempy_b200.engine import compute as gc              # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = ex.synthetic_stress(n_sp_per_surface=1000, n_surfaces=4, n_ori=1000, resolution=(512, 512, 512))
ii, opt, desc = m.args()
eng = gc.B200Engine(0)
pts = ii.grid.dense_grid.n_points // N
out = torch.empty((4, pts), dtype=torch.float64, pin_memory=True)
for aligned in (False, True, False, True):
    for n_slabs in (8, 16):
        ts = []
        for r in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            gc.compute_dense_fields(ii, opt, desc, engine=eng, point_range=(0, pts), out=out, n_slabs=n_slabs, wave_aligned=aligned)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(f"N={N} wave_aligned={aligned} n_slabs={n_slabs}: median {np.median(ts[1:]) * 1e3:.2f} ms  min {min(ts[1:]) * 1e3:.2f} ms", flush=True)
