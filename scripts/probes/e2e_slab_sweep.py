#!/usr/bin/env python
"""Distribution of the compute_dense_fields wall on the 512^3 benchmark for several slab counts (every call printed)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gempy_b200 import examples as ex                    # noqa: E402
from gempy_b200.engine import compute as gc              # noqa: E402

m = ex.synthetic_stress(n_sp_per_surface=1000, n_surfaces=4, n_ori=1000, resolution=(512, 512, 512))
ii, opt, desc = m.args()
eng = gc.B200Engine(0)
pts = ii.grid.dense_grid.n_points
out = torch.empty((4, pts), dtype=torch.float64, pin_memory=True)
for rnd in range(2):
    for n_slabs in (8, 12, 16, 24, 32):
        ts = []
        for r in range(7):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            gc.compute_dense_fields(ii, opt, desc, engine=eng, out=out, n_slabs=n_slabs)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(f"n_slabs={n_slabs}: " + " ".join(f"{t * 1e3:.0f}" for t in ts), flush=True)
