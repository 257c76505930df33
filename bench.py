#!/usr/bin/env python
"""Benchmark of the hot path: grid-point field+gradient evaluations per second (BASELINE.json metric M1).

    python bench.py --gpus N --steps K --warmup W            # our arm (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # restated reference algorithm on host cores

Workload (BASELINE.json configs[2], SURVEY.md 8d "Config 3"): single stack, 4 surfaces x 1000 surface points +
1000 orientations (n = 6999 system), cubic kernel, dense 512^3 regular grid; the grid is sharded by point range
over the N GPUs (strong scaling: the total is fixed at 134 217 728 points).  One step = field + gradient of the
whole grid from resident weights.  `e2e` = the same through the host-facing call with host buffers: H2D of the
input tables, assembly + solve, evaluation, D2H of Z and the gradient into pinned memory.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid_point_field_gradient_evals_per_s"
UNIT = "evals/s"


def flops_per_point(n_ori, n_rest, n_surf, n_drift=3, n_faults=0):
    """BASELINE.md flop model (FMA = 2): field + gradient, cubic kernel."""
    return 41 * n_ori + 32 * (n_rest + n_surf) + (9 if n_drift else 0) + 2 * n_faults


def build_workload(args):
    from gempy_b200 import examples as ex
    res = (args.grid, args.grid, args.grid)
    return ex.synthetic_stress(n_sp_per_surface=args.sp_per_surface, n_surfaces=4, n_ori=args.n_ori, resolution=res)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(payload):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import gempy_oracle as orc
    st, ko, w, xyz = payload
    # the reference's chunking policy: kernel-matrix elements per chunk <= evaluation_chunk_size = 500 000
    Z, G = orc.evaluate(st, ko, w, xyz, gradient=True, chunk_elems=500_000)
    return Z.shape[0]


def cpu_baseline(model, sample_points, cores=None):
    """Restated reference algorithm (numpy float64, oracle/) on the host cores: evaluation of field + gradient on a
    bounded sample of the same grid with the same data, all cores (one process per core over point ranges)."""
    import multiprocessing as mp
    from oracle import gempy_oracle as orc
    ii, opt, desc = model.args()
    ko = opt.kernel_options
    st = orc.prepare_stack(ii.surface_points.sp_coords, ii.surface_points.nugget_effect_scalar,
                           desc.tensors_structure.number_of_points_per_surface, ii.orientations.dip_positions,
                           ii.orientations.dip_gradients, ii.orientations.nugget_effect_grad)
    rng = np.random.default_rng(0)
    w = rng.standard_normal(orc.system_size(st, ko))          # timing only: the weights' values do not matter
    g = ii.grid.dense_grid
    n_total = g.n_points
    idx = np.linspace(0, n_total - 1, sample_points).astype(np.int64)
    s = g.regular_grid_shape
    ix, rem = np.divmod(idx, s[1] * s[2])
    iy, iz = np.divmod(rem, s[2])
    ax = g.axis_coords()
    xyz = np.stack([ax[0][ix], ax[1][iy], ax[2][iz]], axis=1) + orc.GRID_SHIFT
    cores = cores or os.cpu_count() or 1
    parts = np.array_split(xyz, cores)
    t0 = time.perf_counter()
    if cores == 1:
        _cpu_worker((st, ko, w, xyz))
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_worker, [(st, ko, w, p) for p in parts])
    dt = time.perf_counter() - t0
    return sample_points / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = build_workload(args)
    ii, opt, desc = model.args()
    sample = args.cpu_sample
    cores = os.cpu_count() or 1
    times = []
    for _ in range(args.warmup):
        cpu_baseline(model, max(sample // 8, cores * 8), cores)
    for _ in range(args.steps):
        v, dt, cores = cpu_baseline(model, sample, cores)
        times.append(dt)
    dt = float(np.mean(times))
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, model, extra={"timed": f"{sample} grid points sampled uniformly from the {args.grid}^3 grid per step"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of {args.grid ** 3} grid points per step, numpy float64 restatement of the "
                                   "reference algorithm (the engine package is absent from the reference tree), "
                                   "chunked at evaluation_chunk_size=500000 kernel-matrix elements, one process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, model, extra=None):
    ii, opt, desc = model.args()
    n_sp, n_o = ii.surface_points.n_points, ii.orientations.n_items
    n_surf = desc.tensors_structure.n_surfaces
    cfg = {"workload": f"BASELINE configs[2]: synthetic single stack, {n_sp} surface points ({n_surf} surfaces) + {n_o} "
                       f"orientations (n = {3 * n_o + n_sp - n_surf + 3} system), cubic kernel, dense {args.grid}^3 regular grid, "
                       "field + gradient",
           "grid_points": args.grid ** 3, "n_surface_points": n_sp, "n_orientations": n_o,
           "partition": "grid-point ranges (x slabs), no data-path collective",
           "l2_policy": "outputs (4.3 GB per step at 512^3) exceed the 126 MB L2; inputs are a 0.2 MB table"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------ GPU arm
def dfma_peak_tflops(eng):
    """Measured FP64 FMA peak of this GPU: register-resident DFMA chains (gpb_bench_dfma); MEASURED_PEAKS.json has
    no FP64 entry, so the roofline denominator is measured in the same run."""
    import ctypes as C
    from gempy_b200 import _lib
    best = 0.0
    for _ in range(3):
        v = C.c_double(0.0)
        _lib.check(eng.lib.gpb_bench_dfma(20000, C.byref(v), eng.stream))
        best = max(best, v.value)
    return best


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from gempy_b200 import _lib
    from gempy_b200.engine import compute as gc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    eng = gc.B200Engine(local if world > 1 else 0)
    model = build_workload(args)
    ii, opt, desc = model.args()
    ko = opt.kernel_options
    g = ii.grid.dense_grid
    n_total = g.n_points
    # shard: contiguous point ranges, aligned to whole x slabs when possible
    per = (n_total + world - 1) // world
    i0, i1 = rank * per, min(n_total, (rank + 1) * per)
    m = i1 - i0

    # ---- resident state: tables, solved weights, packed evaluation table (every rank solves redundantly)
    st = gc.StackTables(ii, desc, 0, ko, eng.device)
    A, b = eng.assemble(st)
    w = eng.solve(A, b)
    del A
    src = eng.pack(st, w)
    Z = eng.empty(m)
    G = eng.empty(3, m)
    gd = gc.regular_descriptor(g)
    seg = gc.Segment("dense_grid", m, grid=gd, i0=i0)

    def step():
        eng.evaluate_segment(st, src, seg, 0, Z, G, None)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- e2e: host tables in, pinned host arrays out, everything inside the timed region
    e2e_val, h2d, d2h = None, 0, 0
    if not args.no_e2e:
        host_out = torch.empty((4, m), dtype=torch.float64, pin_memory=True)

        def e2e_step():
            # the host-facing call: numpy tables in, pinned host arrays out (H2D, assemble, solve, evaluate, D2H)
            gc.compute_dense_fields(ii, opt, desc, engine=eng, point_range=(i0, i1), out=host_out)
            return float(host_out[0, 0])

        e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        if world > 1:
            dist.barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_val = n_total / float(tt.item())
        h2d = int((ii.surface_points.sp_coords.nbytes + ii.surface_points.nugget_effect_scalar.nbytes +
                   ii.orientations.dip_positions.nbytes + ii.orientations.dip_gradients.nbytes +
                   ii.orientations.nugget_effect_grad.nbytes))
        d2h = int(4 * m * 8)

    if rank == 0:
        n_o, n_rest, n_surf = st.n_ori, st.n_rest, st.n_surf
        F = flops_per_point(n_o, n_rest, n_surf)
        achieved_tf = value / world * F / 1e12                      # per GPU
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        nominal_peak = eng.sm_count * 64 * 2 * 1.965e9 / 1e12       # 148 SM x 64 FP64 lanes x 2 x 1.965 GHz
        peak_at_clock = eng.sm_count * 64 * 2 * sm_mhz * 1e6 / 1e12
        measured_peak = dfma_peak_tflops(eng)
        # DRAM traffic of one launch of the dominant kernel on the default workload, from one `ncu --set full` capture
        # (profiles/r1_eval_zrun_kernel_ncu_full_512.txt): 4.318 GB written + 0.116 GB read vs 4.295 GB algorithmic
        traffic = 4.317651e9 + 0.116064e9 if (args.grid == 512 and world == 1) else None
        roof = {"bound": "fp64", "achieved": achieved_tf, "peak": measured_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / measured_peak, "traffic": traffic, "algorithmic_bytes_per_launch": 32 * m,
                "peak_source": "measured in this run: DFMA-chain microbenchmark gpb_bench_dfma (MEASURED_PEAKS.json has no FP64 entry)",
                "peak_nominal": nominal_peak, "frac_of_nominal": achieved_tf / nominal_peak,
                "peak_at_measured_clock": peak_at_clock, "frac_at_measured_clock": achieved_tf / peak_at_clock,
                "flops_per_point": F, "kernel": "eval_zrun_kernel<cubic, grad, P=8, T=256>",
                "bound_note": "neither HBM (0.06 % DRAM utilisation) nor tensor: the FP64 FMA pipe is the binding resource (ncu: 89 % pipe-active)",
                "hbm": {"achieved_gbs": value / world * 32 / 1e9, "algorithmic_bytes_per_point": 32}}
        cpu = None
        if not args.no_cpu:
            v, dt, cores = cpu_baseline(model, args.cpu_sample)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_sample} of {n_total} grid points ({dt:.1f} s), numpy float64 restatement "
                             "(oracle/), chunked at 500000 kernel-matrix elements, one process per core"}
        # metric M2 (BASELINE.json): compute_model wall seconds on configs[1] (COMBINATION, octree level 6), this GPU
        m2 = None
        if world == 1 and not args.no_m2:
            from gempy_b200 import examples as ex2
            del Z, G
            torch.cuda.empty_cache()
            models = [ex2.combination(refinement=6) for _ in range(6)]        # built outside the timed region
            gc.compute_model(*models[0].args(), engine=eng)
            ts = []
            for mdl in models[1:]:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                gc.compute_model(*mdl.args(), engine=eng)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            m2 = {"model": "COMBINATION octree level 6 (BASELINE configs[1])", "compute_model_wall_s": min(ts),
                  "restated_numpy_wall_s_recorded": 13.1, "record": "profiles/r1_compute_model_wall.jsonl"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, model),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": None if e2e_val is None else {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                                    "d2h_bytes_per_step": d2h},
                "roofline": roof, "cpu_baseline": cpu, "compute_model": m2}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--sp-per-surface", type=int, default=1000)
    ap.add_argument("--n-ori", type=int, default=1000)
    ap.add_argument("--cpu-sample", type=int, default=131072)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-m2", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
