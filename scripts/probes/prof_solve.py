import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from gempy_b200 import _lib, examples as ex
from gempy_b200.engine import compute as gc
eng = gc.B200Engine(0)
n_sp = int(sys.argv[1])
m = ex.synthetic_stress(n_sp_per_surface=n_sp, n_surfaces=4, n_ori=n_sp, resolution=(4, 4, 4))
ii, opt, desc = m.args()
st = gc.StackTables(ii, desc, 0, opt.kernel_options, eng.device); st.set_faults(None)
for _ in range(2):
    w, path = eng.solve_stack(st)
torch.cuda.synchronize()
print(path, float(w.abs().max()))
