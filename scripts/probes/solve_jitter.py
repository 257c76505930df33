"""Wall time and host enqueue time of repeated n = 6999 solves (is the host the bottleneck of the launch-heavy LU?)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gempy_b200 import examples as ex
from gempy_b200.engine import compute as gc
eng = gc.B200Engine(0)
m = ex.synthetic_stress(n_sp_per_surface=1000, n_surfaces=4, n_ori=1000, resolution=(4, 4, 4))
ii, opt, desc = m.args()
st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device)
rows = []
for rep in range(12):
    A, b = eng.assemble(st); torch.cuda.synchronize()
    t = time.perf_counter(); eng.solve(A, b); te = time.perf_counter() - t
    torch.cuda.synchronize(); rows.append((round((time.perf_counter() - t) * 1e3, 1), round(te * 1e3, 1)))
print(os.environ.get("GPB_LU_NO_LOOKAHEAD"), os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), rows)
