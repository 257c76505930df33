#!/usr/bin/env python
"""Measured FP64 peaks of this GPU: DFMA pipe vs DMMA (mma.sync m8n8k4)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gempy_b200 import _lib
from gempy_b200.engine import compute as gc
eng = gc.B200Engine(0)
out = {}
for name in ("gpb_bench_dfma", "gpb_bench_dmma"):
    best = 0.0
    for _ in range(3):
        v = C.c_double(0.0)
        _lib.check(getattr(eng.lib, name)(20000, C.byref(v), eng.stream))
        best = max(best, v.value)
    out[name] = best
for ratio in (4, 8, 16, 32):
    a, b = C.c_double(0.0), C.c_double(0.0)
    _lib.check(eng.lib.gpb_bench_mixed(5000, ratio, C.byref(a), C.byref(b), eng.stream))
    out[f"mixed_ratio_{ratio}"] = {"dfma_tflops": a.value, "dmma_tflops": b.value, "sum": a.value + b.value}
print(json.dumps(out))
