#!/usr/bin/env python
"""BASELINE configs[4]: 20 000 surface points + 5 000 orientations (n ~ 35k FP64 system, 9.8 GB), Matern-5/2,
octree evaluation.  One GPU: assemble, solve, octree levels; reports stage times and self-consistency checks
(no CPU oracle at this size: the residual of the solve and the interpolation conditions stand in)."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import examples as ex
from gempy_b200.engine import compute as gc
from gempy_b200.engine.data import AvailableKernelFunctions as K

ap = argparse.ArgumentParser()
ap.add_argument("--sp-per-surface", type=int, default=5000)
ap.add_argument("--n-ori", type=int, default=5000)
ap.add_argument("--levels", type=int, default=6)
args = ap.parse_args()
eng = gc.B200Engine(0)
m = ex.synthetic_stress(n_sp_per_surface=args.sp_per_surface, n_surfaces=4, n_ori=args.n_ori, kernel=K.matern_5_2,
                        refinement=args.levels)
ii, opt, desc = m.args()
opt.mesh_extraction = False
st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device)
rec = {"n": st.n, "kernel": "matern_5_2", "levels": args.levels}
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, time.perf_counter() - t0
(A, b), rec["assemble_s"] = t(lambda: eng.assemble(st))
A0 = A.clone() if st.n <= 12000 else None
b0 = b.clone()
w, rec["solve_s"] = t(lambda: eng.solve(A, b))
del A
if A0 is not None:
    rec["solve_residual"] = float((A0 @ w - b0).abs().max())
    del A0
else:
    A2, _ = eng.assemble(st)     # re-assemble (the factorisation is in place) and check the residual
    rec["solve_residual"] = float((A2 @ w - b0).abs().max())
    del A2
torch.cuda.empty_cache()
m.interpolation_input.weights = [w.cpu().numpy()]
sol, rec["compute_model_octree_s"] = t(lambda: gc.compute_model(*m.args(), engine=eng))
rec["leaf_counts"] = [int(l.grid_centers.octree_grid.n_points) for l in sol.octrees_output]
# interpolation conditions on a sample of the data
src = eng.pack(st, w)
pts = np.vstack([ii.surface_points.sp_coords[::50], ii.orientations.dip_positions[::50]])
seg = gc.Segment("p", pts.shape[0], xyz=torch.as_tensor(np.ascontiguousarray(pts.T), device=eng.device))
Z = eng.empty(seg.m); G = eng.empty(3, seg.m)
eng.evaluate_segment(st, src, seg, 0, Z, G, None)
Zh, Gh = Z.cpu().numpy(), G.cpu().numpy().T
nsp = ii.surface_points.sp_coords[::50].shape[0]
per = args.sp_per_surface // 50
rec["max_dev_Z_on_surface"] = float(max(np.abs(Zh[k*per:(k+1)*per] - Zh[k*per:(k+1)*per].mean()).max() for k in range(4)))
rec["max_dev_grad_at_orientations"] = float(np.abs(Gh[nsp:] - ii.orientations.dip_gradients[::50]).max())
print(json.dumps(rec))
