import sys, os, time
import torch
sys.path.insert(0, os.getcwd())
from gempy_b200 import examples as ex
from gempy_b200.engine import compute as gc
eng = gc.B200Engine(0)
build = lambda: ex.synthetic_multi_fault(refinement=8)
def stats():
    s = torch.cuda.memory_stats()
    return s.get("num_device_alloc", 0), s.get("num_device_free", 0), s.get("num_alloc_retries", 0), s["reserved_bytes.all.current"] >> 20, s["allocated_bytes.all.current"] >> 20
sol = None
for it in range(6):
    m = build(); torch.cuda.synchronize(); a = stats(); t0 = time.perf_counter()
    sol = gc.compute_model(*m.args(), engine=eng)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0; b = stats()
    print(f"iter {it}: {dt*1e3:7.1f} ms  cudaMalloc +{b[0]-a[0]} cudaFree +{b[1]-a[1]} retries +{b[2]-a[2]} reserved {b[3]} MiB allocated {b[4]} MiB", flush=True)
print("-- dropping the previous solution before each call")
for it in range(4):
    m = build(); sol = None; torch.cuda.synchronize(); a = stats(); t0 = time.perf_counter()
    sol = gc.compute_model(*m.args(), engine=eng)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0; b = stats()
    print(f"iter {it}: {dt*1e3:7.1f} ms  cudaMalloc +{b[0]-a[0]} cudaFree +{b[1]-a[1]} retries +{b[2]-a[2]} reserved {b[3]} MiB allocated {b[4]} MiB", flush=True)
import gc as _gc
print("gc garbage check: collected", _gc.collect())
