#!/usr/bin/env python
"""Time the blocked LU (gpb_lu_solve) against cuSOLVER getrf/getrs (through torch.linalg) on the co-kriging
system of the benchmark model (BASELINE north_star: "timed against cuSOLVER getrf/getrs")."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import examples as ex                 # noqa: E402
from gempy_b200.engine import compute as gc           # noqa: E402


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.min(ts)), float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sp-per-surface", type=int, default=1000)
    ap.add_argument("--n-ori", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--outer-min-n", type=int, default=-1, help="force the outer-blocked schedule from this n on")
    ap.add_argument("--skip-cusolver", action="store_true")
    ap.add_argument("--outer-width", type=int, default=0)
    args = ap.parse_args()
    eng = gc.B200Engine(0)
    eng.lib.gpb_lu_set_outer_width(args.outer_width)
    if args.outer_min_n >= 0:
        eng.lib.gpb_lu_set_outer_min_n(args.outer_min_n)
    m = ex.synthetic_stress(n_sp_per_surface=args.sp_per_surface, n_surfaces=4, n_ori=args.n_ori, resolution=(4, 4, 4))
    ii, opt, desc = m.args()
    st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device)
    A0, b0 = eng.assemble(st)
    n = A0.shape[0]
    torch.cuda.synchronize()
    out = {}

    def ours():
        A, b = A0.clone(), b0.clone()
        out["w"] = eng.solve(A, b)

    def cusolver():
        LU, piv = torch.linalg.lu_factor(A0)          # A0 is symmetric: row/column-major agree
        out["w_ref"] = torch.linalg.lu_solve(LU, piv, b0[:, None])[:, 0]

    def clone_only():
        A0.clone(); b0.clone()

    if args.skip_cusolver:
        def cusolver():
            out["w_ref"] = out["w"]
    for f in (ours, cusolver, clone_only):
        f()
    t_ours = timed(ours, args.reps)
    t_cus = timed(cusolver, args.reps)
    t_clone = timed(clone_only, args.reps)
    w, w_ref = out["w"], out["w_ref"]
    r = (A0 @ w - b0).abs().max().item()
    r_ref = (A0 @ w_ref - b0).abs().max().item()
    flops = 2.0 / 3.0 * n ** 3
    print(json.dumps({"n": n, "outer_min_n": int(eng.lib.gpb_lu_set_outer_min_n(-1)), "gpb_lu_solve_ms": t_ours[0] - t_clone[0], "cusolver_getrf_getrs_ms": t_cus[0],
                      "clone_ms": t_clone[0], "gpb_tflops": flops / ((t_ours[0] - t_clone[0]) * 1e-3) / 1e12,
                      "cusolver_tflops": flops / (t_cus[0] * 1e-3) / 1e12, "residual_ours": r, "residual_cusolver": r_ref,
                      "max_rel_diff_weights": ((w - w_ref).abs().max() / w_ref.abs().max()).item()}))


if __name__ == "__main__":
    main()
