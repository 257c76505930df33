#!/usr/bin/env python
"""HBM-bound stages (activator, combine, covariance assembly) against the measured copy bandwidth
(MEASURED_PEAKS.json hbm_gbs)."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import _lib, examples as ex
from gempy_b200.engine import compute as gc

eng = gc.B200Engine(0)
peak = 6550.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

def timed(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

m = 1 << 27                                     # 134 M points (512^3)
Z = torch.rand(m, dtype=torch.float64, device=eng.device)
block = torch.empty_like(Z)
iso = torch.tensor([0.8, 0.6, 0.4, 0.2], dtype=torch.float64, device=eng.device)
ids = torch.tensor([1., 2., 3., 4., 5.], dtype=torch.float64, device=eng.device)
out = {"hbm_peak_gbs": peak, "points": m}
ms = timed(lambda: _lib.check(eng.lib.gpb_activate(Z.data_ptr(), m, iso.data_ptr(), ids.data_ptr(), 4, 5e6, block.data_ptr(), eng.stream)))
out["activate"] = {"ms": ms, "algorithmic_bytes": 16 * m, "gbs": 16 * m / ms / 1e6, "frac": 16 * m / ms / 1e6 / peak}
n_st = 3
Zs = torch.rand(n_st, m, dtype=torch.float64, device=eng.device)
Bs = torch.rand(n_st, m, dtype=torch.float64, device=eng.device)
fb, fa = torch.empty(m, dtype=torch.float64, device=eng.device), torch.empty(m, dtype=torch.float64, device=eng.device)
sq = torch.empty(n_st, m, dtype=torch.uint8, device=eng.device); mk = torch.empty(n_st, m, dtype=torch.uint8, device=eng.device)
imin = torch.tensor([0.3, 0.3, 0.3], dtype=torch.float64, device=eng.device); imax = imin + 0.3
rel = (C.c_int * 3)(3, 1, 4)
ms = timed(lambda: _lib.check(eng.lib.gpb_combine(Zs.data_ptr(), Bs.data_ptr(), m, m, n_st, rel, imin.data_ptr(), imax.data_ptr(), fb.data_ptr(), fa.data_ptr(), sq.data_ptr(), mk.data_ptr(), eng.stream)))
byts = (2 * n_st * 8 + 16 + 2 * n_st) * m
out["combine_3_stacks"] = {"ms": ms, "algorithmic_bytes": byts, "gbs": byts / ms / 1e6, "frac": byts / ms / 1e6 / peak}
mdl = ex.synthetic_stress(n_sp_per_surface=1000, n_surfaces=4, n_ori=1000, resolution=(4, 4, 4))
ii, opt, desc = mdl.args()
st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device)
n = st.n
A = eng.empty(n, n); b = eng.empty(n); s = st.struct()
ms = timed(lambda: _lib.check(eng.lib.gpb_assemble_cov(C.byref(s), A.data_ptr(), n, b.data_ptr(), eng.stream)))
out["assemble_cov_n6999"] = {"ms": ms, "algorithmic_bytes": 8 * n * n, "gbs": 8 * n * n / ms / 1e6, "frac": 8 * n * n / ms / 1e6 / peak}
print(json.dumps(out))
