#!/usr/bin/env python
"""Solve of the co-kriging saddle-point system: symmetric path (gpb_sym_solve: Cholesky of the covariance block +
Schur complement) and pivoted LU (gpb_lu_solve) against cuSOLVER (getrf + getrs, and potrf + potrs on the covariance
block) through torch.linalg, same process, same matrix (BASELINE north_star: "timed against cuSOLVER getrf/getrs").

    python scripts/bench_solve.py --sizes 250,1000,2500 [--kernel matern_5_2] [--reps 5]
`--sizes` are surface points per surface (4 surfaces) with as many orientations: n = 3 n_o + 4 n_sp - 4 + 3."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import _lib, examples as ex           # noqa: E402
from gempy_b200.engine import compute as gc           # noqa: E402
from gempy_b200.engine.data import AvailableKernelFunctions as K   # noqa: E402


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


def run(eng, n_sp, n_ori, kernel, reps, skip_cusolver=False):
    m = ex.synthetic_stress(n_sp_per_surface=n_sp, n_surfaces=4, n_ori=n_ori, resolution=(4, 4, 4))
    ii, opt, desc = m.args()
    opt.kernel_options.kernel_function = kernel
    st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device)
    st.set_faults(None)
    A0, b0 = eng.assemble(st)                        # full symmetric n x n
    n = A0.shape[0]
    nk = 3 * st.n_ori + st.n_rest
    lda = (n + 2) & ~1
    Apad = eng.empty(n, lda)
    out = {}
    info = torch.zeros(1, dtype=torch.int32, device=eng.device)

    def sym():
        Apad[:, :n].copy_(A0)
        b = b0.clone()
        _lib.check(eng.lib.gpb_sym_solve(n, nk, Apad.data_ptr(), lda, b.data_ptr(), 1, n, info.data_ptr(), eng.stream))
        out["w_sym"] = b

    def lu():
        Apad[:, :n].copy_(A0)
        b = b0.clone()
        ipiv = eng.empty(n, dtype=torch.int32)
        _lib.check(eng.lib.gpb_lu_solve(n, Apad.data_ptr(), lda, b.data_ptr(), 1, n, ipiv.data_ptr(), info.data_ptr(), eng.stream))
        out["w_lu"] = b

    def copy_only():
        Apad[:, :n].copy_(A0)
        b0.clone()

    def cus_getrf():
        LU, piv = torch.linalg.lu_factor(A0)         # A0 is symmetric: row/column-major agree
        out["w_getrf"] = torch.linalg.lu_solve(LU, piv, b0[:, None])[:, 0]

    Kb = A0[:nk, :nk].contiguous()

    def cus_potrf():                                 # the comparable library call for the covariance block alone
        L = torch.linalg.cholesky(Kb)
        out["y_potrf"] = torch.cholesky_solve(b0[:nk, None], L)

    def clone_full():
        A0.clone()

    sct = st.struct()

    def cov_full():
        _lib.check(eng.lib.gpb_assemble_cov_ex(ctypes.byref(sct), Apad.data_ptr(), lda, b0.data_ptr(), 0, eng.stream))

    def cov_lower():
        _lib.check(eng.lib.gpb_assemble_cov_ex(ctypes.byref(sct), Apad.data_ptr(), lda, b0.data_ptr(), 1, eng.stream))

    fns = [sym, lu, copy_only, clone_full, cov_full, cov_lower] + ([] if skip_cusolver else [cus_getrf, cus_potrf])
    for f in fns:
        f()
    torch.cuda.synchronize()
    info_sym = None
    sym(); info_sym = int(info.item())
    t = {f.__name__: timed(f, reps) for f in fns}
    w_sym, w_lu = out["w_sym"], out["w_lu"]
    w_ref = out.get("w_getrf", w_lu)
    res = lambda w: float((A0 @ w - b0).abs().max().item())
    rec = {"n": n, "nk": nk, "kernel": kernel.name, "info_sym": info_sym,
           "gpb_sym_solve_ms": t["sym"][0] - t["copy_only"][0], "gpb_lu_solve_ms": t["lu"][0] - t["copy_only"][0],
           "copy_ms": t["copy_only"][0],
           "residual_sym": res(w_sym), "residual_lu": res(w_lu), "residual_cusolver": res(w_ref),
           "rel_diff_sym_vs_cusolver": float(((w_sym - w_ref).abs().max() / w_ref.abs().max()).item()),
           "rel_diff_lu_vs_cusolver": float(((w_lu - w_ref).abs().max() / w_ref.abs().max()).item())}
    if not skip_cusolver:
        # torch.linalg.lu_factor / cholesky clone their input internally; subtract the same clone we subtract from ours
        rec["cusolver_getrf_getrs_ms"] = t["cus_getrf"][0] - t["clone_full"][0]
        rec["cusolver_potrf_potrs_ms_cov_block_only"] = t["cus_potrf"][0] - t["clone_full"][0]
    hbm = 6550.4e9
    rec["cov_assemble_full_ms"] = t["cov_full"][0]
    rec["cov_assemble_lower_ms"] = t["cov_lower"][0]
    rec["cov_full_gbs"] = 8.0 * n * n / (t["cov_full"][0] * 1e-3) / 1e9
    rec["cov_full_frac_of_measured_hbm"] = 8.0 * n * n / (t["cov_full"][0] * 1e-3) / hbm
    rec["cov_lower_frac_of_measured_hbm"] = 4.0 * n * (n + 1) / (t["cov_lower"][0] * 1e-3) / hbm
    rec["sym_tflops_chol_model"] = (nk ** 3 / 3.0) / (rec["gpb_sym_solve_ms"] * 1e-3) / 1e12
    rec["lu_tflops"] = (2.0 * n ** 3 / 3.0) / (rec["gpb_lu_solve_ms"] * 1e-3) / 1e12
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="60,250,1000")
    ap.add_argument("--kernel", default="cubic")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-cusolver", action="store_true")
    args = ap.parse_args()
    eng = gc.B200Engine(0)
    for s in args.sizes.split(","):
        n_sp = int(s)
        rec = run(eng, n_sp, n_sp, K[args.kernel], args.reps, args.skip_cusolver)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
