import sys, os, time, json
import torch
sys.path.insert(0, os.getcwd())
from gempy_b200 import _lib, examples as ex
from gempy_b200.engine import compute as gc
eng = gc.B200Engine(0)
build = lambda: ex.synthetic_multi_fault(refinement=8)
gc.compute_model(*build().args(), engine=eng)
walls=[]
for it in range(4):
    m = build(); torch.cuda.synchronize(); t0=time.perf_counter(); sol = gc.compute_model(*m.args(), engine=eng); torch.cuda.synchronize(); walls.append(time.perf_counter()-t0)
print("walls", [round(w*1e3,1) for w in walls])
# instrumented
T = {}
def wrap(obj, name):
    fn = getattr(obj, name)
    def w(*a, **k):
        torch.cuda.synchronize(); t0=time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize(); T[name] = T.get(name,0)+time.perf_counter()-t0; return r
    setattr(obj, name, w)
for n in ("run_level","emit_into","count_marked","mark","copy_rows","corners_into","corners_of","gather_fields","empty"):
    wrap(eng, n)
orig_dc = gc._dual_contouring
def dc(*a, **k):
    torch.cuda.synchronize(); t0=time.perf_counter(); r = orig_dc(*a, **k); torch.cuda.synchronize(); T["dual_contouring"]=time.perf_counter()-t0; return r
gc._dual_contouring = dc
orig_mt = gc.ModelTables
m = build(); torch.cuda.synchronize(); t0=time.perf_counter(); sol = gc.compute_model(*m.args(), engine=eng); torch.cuda.synchronize(); tot=time.perf_counter()-t0
print("instrumented total ms", round(tot*1e3,1), {k: round(v*1e3,2) for k,v in T.items()})
