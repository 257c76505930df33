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
del A, Zs, Bs, fb, fa, sq, mk, block
# ---- marching cubes on a 512^3 lattice (blob surface), pass by pass
nn = 512
ax = torch.linspace(-1, 1, nn, dtype=torch.float64, device=eng.device)
F = (0.7 - torch.sqrt(ax[:, None, None] ** 2 + 1.3 * ax[None, :, None] ** 2 + 0.8 * ax[None, None, :] ** 2)).contiguous().view(-1)
msk = torch.ones(m, dtype=torch.uint8, device=eng.device)
flags = torch.empty(m, dtype=torch.uint8, device=eng.device)
offs = torch.empty(int(eng.lib.gpb_mc_scratch_elems(m)), dtype=torch.int64, device=eng.device)
nv, nt = C.c_longlong(), C.c_longlong()
count = lambda: _lib.check(eng.lib.gpb_mc_count(F.data_ptr(), msk.data_ptr(), nn, nn, nn, 0.0, flags.data_ptr(), offs.data_ptr(), C.byref(nv), C.byref(nt), eng.stream))
ms_c = timed(count)
vb = torch.empty(m, dtype=torch.int32, device=eng.device)
V = torch.empty(nv.value, 3, dtype=torch.float64, device=eng.device); T = torch.empty(nt.value, 3, dtype=torch.int32, device=eng.device)
emit = lambda: _lib.check(eng.lib.gpb_mc_emit(F.data_ptr(), flags.data_ptr(), offs.data_ptr(), nn, nn, nn, 0.0, 0., 0., 0., 1., 1., 1., vb.data_ptr(), V.data_ptr(), T.data_ptr(), eng.stream))
ms_e = timed(emit)
b_c = 10 * m
b_e = 6 * m + 24 * nv.value + 12 * nt.value
out["marching_cubes_512"] = {"vertices": nv.value, "triangles": nt.value,
                             "count": {"ms": ms_c, "algorithmic_bytes": b_c, "gbs": b_c / ms_c / 1e6, "frac": b_c / ms_c / 1e6 / peak,
                                       "note": "classify + scan + host read of the totals; Z 8 B + mask 1 B read, flags 1 B written per point"},
                             "emit": {"ms": ms_e, "algorithmic_bytes": b_e, "gbs": b_e / ms_e / 1e6, "frac": b_e / ms_e / 1e6 / peak,
                                      "note": "vertices + triangles passes; flags read twice (2 B), vbase 4 B written per point, plus the mesh"}}
print(json.dumps(out))
