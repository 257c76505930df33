#!/usr/bin/env python
"""Where do the sporadic slow end-to-end steps go?  compute_dense_fields re-enacted with CUDA events after the solve, after
every slab's kernel and after every slab's read-back; slow calls print their timeline."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gempy_b200 import examples as ex                    # noqa: E402
from gempy_b200.engine import compute as gc              # noqa: E402
from gempy_b200 import _lib                              # noqa: E402

m = ex.synthetic_stress(n_sp_per_surface=1000, n_surfaces=4, n_ori=1000, resolution=(512, 512, 512))
ii, opt, desc = m.args()
eng = gc.B200Engine(0)
g = ii.grid.dense_grid
pts = g.n_points
out = torch.empty((4, pts), dtype=torch.float64, pin_memory=True)
n_slabs = 16
ev = lambda: torch.cuda.Event(enable_timing=True)
for call in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with eng.hold_stream():
        compute = torch.cuda.current_stream(eng.device)
        e_start = ev(); e_start.record(compute)
        st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device)
        t_tab = time.perf_counter()
        ta = time.perf_counter()
        A, b = eng.assemble(st, extra_rows=1, lower_only=True)
        tb = time.perf_counter()
        info = torch.zeros(1, dtype=torch.int32, device=eng.device)
        nk = 3 * st.n_ori + st.n_rest
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.gpb_sym_solve(st.n, nk, A.data_ptr(), A.shape[1], b.data_ptr(), 1, st.n, info.data_ptr(), eng.stream))
        tc = time.perf_counter()
        assert int(info.item()) == 0
        td = time.perf_counter()
        w = b
        del A
        t_solve_host = time.perf_counter()
        solve_parts = f"[assemble-enqueue {(tb - ta) * 1e3:.1f} sym_solve-enqueue {(tc - tb) * 1e3:.1f} info.item {(td - tc) * 1e3:.1f} free {(t_solve_host - td) * 1e3:.1f}]"
        src = eng.pack(st, w)
        e_solved = ev(); e_solved.record(compute)
        gd = gc.regular_descriptor(g)
        wave = 148 * 256 * 8
        per = -(-(-(-pts // n_slabs)) // wave) * wave
        if getattr(eng, "_copier", None) is None:
            eng._copier = torch.cuda.Stream(eng.device)
        copier = eng._copier
        copier.wait_stream(compute)
        bufs = [eng.empty(4, per) for _ in range(2)]
        free_ev = [None, None]
        k = 0
        marks = []
        for s0 in range(0, pts, per):
            s1 = min(pts, s0 + per)
            buf = bufs[k & 1]
            if free_ev[k & 1] is not None:
                compute.wait_event(free_ev[k & 1])
            seg = gc.Segment("dense_grid", s1 - s0, grid=gd, i0=s0)
            eng.evaluate_segment(st, src, seg, 0, buf[0], buf[1:], None)
            done = ev(); done.record(compute)
            with torch.cuda.stream(copier):
                copier.wait_event(done)
                for a in range(4):
                    out[a, s0:s1].copy_(buf[a, :s1 - s0], non_blocking=True)
                e = ev(); e.record(copier)
                free_ev[k & 1] = e
            marks.append((done, e))
            k += 1
        t_enq = time.perf_counter()
        copier.synchronize()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    wall = (t1 - t0) * 1e3
    solved = e_start.elapsed_time(e_solved)
    ker = [e_start.elapsed_time(d) for d, _ in marks]
    cop = [e_start.elapsed_time(c) for _, c in marks]
    dk = np.diff([solved] + ker)
    lag = [c - d for d, c in zip(ker, cop)]
    print(f"call {call}: wall {wall:.0f} ms | host: tables {(t_tab - t0) * 1e3:.1f} solve-enqueue {(t_solve_host - t_tab) * 1e3:.1f} "
          f"{solve_parts} enqueue-done {(t_enq - t0) * 1e3:.1f} | device: solved at {solved:.1f}, kernels end {ker[-1]:.0f}, copies end {cop[-1]:.0f} | "
          f"slab kernel max {dk.max():.1f} min {dk.min():.1f} | copy lag max {max(lag):.1f} min {min(lag):.1f}", flush=True)
    if wall > 900:
        print("   slab kernel ms:", " ".join(f"{x:.0f}" for x in dk))
        print("   copy lag ms:   ", " ".join(f"{x:.0f}" for x in lag), flush=True)
