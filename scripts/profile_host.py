import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gempy_b200 import examples as ex
from gempy_b200.engine import compute as gc
eng = gc.B200Engine(0)
for name, build in (("combination6", lambda: ex.combination(refinement=6)), ("multifault8", lambda: ex.synthetic_multi_fault(refinement=8))):
    gc.compute_model(*build().args(), engine=eng)
    m = build()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    gc.compute_model(*m.args(), engine=eng)
    torch.cuda.synchronize()
    pr.disable()
    print(name, "wall", time.perf_counter() - t0)
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(s.getvalue()[:6000])
