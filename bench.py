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
    st, ko, w, xyz, chunk = payload
    Z, G = orc.evaluate(st, ko, w, xyz, gradient=True, chunk_elems=chunk)
    return Z.shape[0]


def _cpu_sample(model, sample_points):
    """The oracle-side stack tables, a weight vector (timing only: values do not matter) and `sample_points` grid points
    spread uniformly over the dense grid."""
    from oracle import gempy_oracle as orc
    ii, opt, desc = model.args()
    ko = opt.kernel_options
    st = orc.prepare_stack(ii.surface_points.sp_coords, ii.surface_points.nugget_effect_scalar,
                           desc.tensors_structure.number_of_points_per_surface, ii.orientations.dip_positions,
                           ii.orientations.dip_gradients, ii.orientations.nugget_effect_grad)
    rng = np.random.default_rng(0)
    w = rng.standard_normal(orc.system_size(st, ko))
    g = ii.grid.dense_grid
    idx = np.linspace(0, g.n_points - 1, sample_points).astype(np.int64)
    s = g.regular_grid_shape
    ix, rem = np.divmod(idx, s[1] * s[2])
    iy, iz = np.divmod(rem, s[2])
    ax = g.axis_coords()
    xyz = np.stack([ax[0][ix], ax[1][iy], ax[2][iz]], axis=1) + orc.GRID_SHIFT
    return st, ko, w, xyz


def cpu_baseline(model, sample_points, cores=None, backend="numpy", chunk_elems=500_000):
    """Restated reference algorithm (oracle/) on the host cores: field + gradient on a bounded sample of the same grid
    with the same data.  backend "numpy": one process per core over point ranges (numpy float64); backend "torch": one
    process, torch float64 CPU tensors with `cores` intra-op threads (the reference's PYTORCH backend on CPU).
    chunk_elems = kernel-matrix elements per chunk (the reference's evaluation_chunk_size, default 500 000)."""
    import multiprocessing as mp
    st, ko, w, xyz = _cpu_sample(model, sample_points)
    cores = cores or os.cpu_count() or 1
    if backend == "torch":
        import torch
        from oracle import gempy_oracle_torch as ot
        torch.set_num_threads(cores)
        t0 = time.perf_counter()
        ot.evaluate(st, ko, w, xyz, gradient=True, chunk_elems=chunk_elems)
        dt = time.perf_counter() - t0
        return sample_points / dt, dt, cores
    parts = np.array_split(xyz, cores)
    t0 = time.perf_counter()
    if cores == 1:
        _cpu_worker((st, ko, w, xyz, chunk_elems))
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_worker, [(st, ko, w, p, chunk_elems) for p in parts])
    dt = time.perf_counter() - t0
    return sample_points / dt, dt, cores


def _cpu_desc(backend, chunk):
    if backend == "torch":
        return (f"PyTorch-CPU float64 restatement of the reference algorithm (oracle/gempy_oracle_torch.py; the engine package "
                f"is absent from the reference tree), chunked at {chunk} kernel-matrix elements, one process, all cores as intra-op threads")
    return (f"numpy float64 restatement of the reference algorithm (oracle/gempy_oracle.py; the engine package is absent from "
            f"the reference tree), chunked at {chunk} kernel-matrix elements, one process per core")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = build_workload(args)
    sample = args.cpu_sample
    cores = os.cpu_count() or 1
    times = []
    for _ in range(args.warmup):
        cpu_baseline(model, max(sample // 8, cores * 8), cores, args.backend, args.cpu_chunk)
    for _ in range(args.steps):
        v, dt, cores = cpu_baseline(model, sample, cores, args.backend, args.cpu_chunk)
        times.append(dt)
    dt = float(np.mean(times))
    value = sample / dt
    line = {
        "impl": "reference", "backend": args.backend, "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, model, extra={"timed": f"{sample} grid points sampled uniformly from the {args.grid}^3 grid per step"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of {args.grid ** 3} grid points per step, " + _cpu_desc(args.backend, args.cpu_chunk)},
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
def fp64_peaks(eng):
    """Measured FP64 peaks of this GPU, in this process: register-resident DFMA chains (gpb_bench_dfma) and DMMA m8n8k4
    chains (gpb_bench_dmma).  MEASURED_PEAKS.json has no FP64 entry, so the roofline denominator is measured next to the
    number it divides (same clocks, same box)."""
    import ctypes as C
    from gempy_b200 import _lib
    out = {}
    for name, fn in (("dfma_tflops", eng.lib.gpb_bench_dfma), ("dmma_tflops", eng.lib.gpb_bench_dmma)):
        best = 0.0
        for _ in range(3):
            v = C.c_double(0.0)
            _lib.check(fn(100000, C.byref(v), eng.stream))
            best = max(best, v.value)
        out[name] = best
    return out


def _timed_ms(torch, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.min(ts))


def solve_record(torch, gc, _lib, eng, st, reps=5):
    """The n = 6999 saddle-point solve of the benchmark model: symmetric path (gpb_sym_solve), pivoted LU (gpb_lu_solve)
    and cuSOLVER getrf + getrs / potrf + potrs through torch.linalg, same process, same matrix; copies of the input are
    timed separately and subtracted from every arm (torch.linalg clones internally)."""
    A0, b0 = eng.assemble(st)
    n = A0.shape[0]
    nk = 3 * st.n_ori + st.n_rest
    lda = (n + 2) & ~1
    Apad = eng.empty(n, lda)
    info = torch.zeros(1, dtype=torch.int32, device=eng.device)
    ipiv = eng.empty(n, dtype=torch.int32)
    out = {}

    def sym():
        Apad[:, :n].copy_(A0)
        b = b0.clone()
        _lib.check(eng.lib.gpb_sym_solve(n, nk, Apad.data_ptr(), lda, b.data_ptr(), 1, n, info.data_ptr(), eng.stream))
        out["sym"] = b

    def lu():
        Apad[:, :n].copy_(A0)
        b = b0.clone()
        _lib.check(eng.lib.gpb_lu_solve(n, Apad.data_ptr(), lda, b.data_ptr(), 1, n, ipiv.data_ptr(), info.data_ptr(), eng.stream))
        out["lu"] = b

    def copy_only():
        Apad[:, :n].copy_(A0)
        b0.clone()

    def getrf():
        LU, piv = torch.linalg.lu_factor(A0)
        out["getrf"] = torch.linalg.lu_solve(LU, piv, b0[:, None])[:, 0]

    Kb = A0[:nk, :nk].contiguous()

    def potrf():
        L = torch.linalg.cholesky(Kb)
        out["potrf"] = torch.cholesky_solve(b0[:nk, None], L)

    def clone_full():
        A0.clone()

    for f in (sym, lu, copy_only, getrf, potrf, clone_full):
        f()
    t = {f.__name__: _timed_ms(torch, f, reps) for f in (sym, lu, copy_only, getrf, potrf, clone_full)}
    res = lambda w: float((A0 @ w - b0).abs().max().item())
    ref = out["getrf"]
    return {"n": n, "gpb_sym_solve_ms": t["sym"] - t["copy_only"], "gpb_lu_solve_ms": t["lu"] - t["copy_only"],
            "cusolver_getrf_getrs_ms": t["getrf"] - t["clone_full"],
            "cusolver_potrf_potrs_ms_covariance_block_only": t["potrf"] - t["clone_full"],
            "residual_sym": res(out["sym"]), "residual_lu": res(out["lu"]), "residual_cusolver": res(ref),
            "rel_diff_sym_vs_cusolver": float(((out["sym"] - ref).abs().max() / ref.abs().max()).item()),
            "note": "same process, same matrix; min of %d runs each; input copies subtracted from every arm" % reps}


def compute_model_record(torch, gc, eng, comm, build, reads, reps=3):
    """compute_model wall (metric M2): host tables in, Solutions out, then the arrays a user reads are pulled to the
    host (`reads`: a function of the Solutions touching them).  Returns the wall to the lazy Solutions and the wall
    including the host reads; max over ranks is taken by the caller."""
    walls, walls_read = [], []
    info = {}
    for r in range(reps + 1):
        m = build()                                   # built outside the timed region
        torch.cuda.synchronize()
        comm_barrier(comm)
        t0 = time.perf_counter()
        sol = gc.compute_model(*m.args(), engine=eng, comm=comm)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        info = reads(sol)
        t2 = time.perf_counter()
        if r > 0:                                     # first run: warm-up (allocator, module load)
            walls.append(t1 - t0)
            walls_read.append(t2 - t0)
        info["levels"] = [int(l.grid_centers.octree_grid.n_points) for l in sol.octrees_output]
        del sol
    return min(walls), min(walls_read), info


def comm_barrier(comm):
    if comm is not None and comm.world > 1:
        import torch.distributed as dist
        dist.barrier()


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from gempy_b200 import _lib
    from gempy_b200 import examples as ex
    from gempy_b200.engine import compute as gc
    from gempy_b200.engine.comm import Comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    eng = gc.B200Engine(local if world > 1 else 0)
    comm = Comm() if world > 1 else None
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
    w, solver_path = eng.solve_stack(st)
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

    # ---- e2e: host tables in, pinned host arrays out, everything inside the timed region, every step
    e2e_val, h2d, d2h = None, 0, 0
    if not args.no_e2e:
        host_out = torch.empty((4, m), dtype=torch.float64, pin_memory=True)

        def e2e_step():
            # the host-facing call: numpy tables in, pinned host arrays out (H2D, assemble, solve, evaluate, D2H)
            gc.compute_dense_fields(ii, opt, desc, engine=eng, point_range=(i0, i1), out=host_out)
            return float(host_out[0, 0])

        e2e_step()
        # every step timed on its own clock on every rank (one barrier before the first step, none in between: the ranks
        # share nothing on this path), then the per-step MAX over ranks: the shared hosts stall for tens to hundreds of
        # milliseconds now and then (scripts/probes/alloc_trace.py shows the same outliers on plain compute_model calls), so
        # the value is taken from the MEDIAN step; mean and max are reported next to it
        n_e2e = max(1, args.steps)
        if world > 1:
            dist.barrier()
        local_s = []
        for _ in range(n_e2e):
            t0 = time.perf_counter()
            e2e_step()
            local_s.append(time.perf_counter() - t0)
        tt = torch.tensor(local_s, dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_s = [float(v) for v in tt.cpu().numpy()]
        e2e_val = n_total / float(np.median(step_s))
        h2d = int((ii.surface_points.sp_coords.nbytes + ii.surface_points.nugget_effect_scalar.nbytes +
                   ii.orientations.dip_positions.nbytes + ii.orientations.dip_gradients.nbytes +
                   ii.orientations.nugget_effect_grad.nbytes))
        d2h = int(4 * m * 8)
        del host_out
    del Z, G
    torch.cuda.empty_cache()

    # ---- extra records witnessed by the driver at every N ("next round" 1b)
    extras = {}
    if not args.no_extras:
        def mx(v):
            tv = torch.tensor([v], dtype=torch.float64, device=eng.device)
            if world > 1:
                dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            return float(tv.item())

        # (a) the solve, on rank 0 (the other ranks wait at the barrier)
        if rank == 0:
            extras["solve"] = solve_record(torch, gc, _lib, eng, st)
            extras["solve"]["path_used_by_the_bench"] = solver_path
        comm_barrier(comm)
        torch.cuda.empty_cache()

        # (b) configs[2] through compute_model itself: dense 512^3, all outputs of the engine kept on the device, the
        #     lithology block and the scalar field read on the host
        def reads3(sol):
            lb = sol.raw_arrays.lith_block
            sf = sol.raw_arrays.scalar_field_matrix
            return {"lith_block_points": int(lb.shape[0]), "scalar_field_matrix": list(sf.shape),
                    "n_units": int(np.unique(lb[:: max(1, lb.shape[0] // 65536) | 1]).shape[0])}

        def build3():
            m3 = build_workload(args)
            m3.options.evaluation_options.compute_scalar_gradient = True      # the metric's work: field AND gradient
            return m3

        w3, w3r, info3 = compute_model_record(torch, gc, eng, comm, build3, reads3, reps=2)
        extras["compute_model_cfg3"] = {"model": f"BASELINE configs[2] through compute_model: dense {args.grid}^3, scalar field + gradient + activator + lithology block, all kept on the device",
                                        "wall_s": mx(w3), "wall_with_host_reads_s": mx(w3r), **info3,
                                        "host_reads": "raw_arrays.lith_block and raw_arrays.scalar_field_matrix (pageable numpy arrays)"}
        torch.cuda.empty_cache()

        # (c) configs[3]: multi-fault model (10 fault stacks + 5 series with fault drift), octree level 8, dual contouring
        def reads4(sol):
            lb = sol.raw_arrays.lith_block
            nv = sum(int(mm.vertices.shape[0]) for mm in sol.dc_meshes)
            nt = sum(int(mm.edges.shape[0]) for mm in sol.dc_meshes)
            return {"lith_block_points": int(lb.shape[0]), "n_meshes": len(sol.dc_meshes), "n_vertices": nv, "n_triangles": nt}

        w4, w4r, info4 = compute_model_record(torch, gc, eng, comm, lambda: ex.synthetic_multi_fault(refinement=args.cfg4_levels), reads4, reps=3)
        extras["compute_model_cfg4"] = {"model": f"BASELINE configs[3]: 10 fault stacks + 5 series (15 stacks, fault drift), octree level {args.cfg4_levels}, dual contouring",
                                        "wall_s": mx(w4), "wall_with_host_reads_s": mx(w4r), **info4,
                                        "host_reads": "raw_arrays.lith_block (octree -> regular fill on the device) and every mesh's vertices / edges"}
        # (d) configs[1]: COMBINATION at octree level 6
        if world == 1:
            def reads2(sol):
                return {"lith_block_points": int(sol.raw_arrays.lith_block.shape[0]), "n_meshes": len(sol.dc_meshes)}
            w2, w2r, info2 = compute_model_record(torch, gc, eng, comm, lambda: ex.combination(refinement=6), reads2, reps=5)
            extras["compute_model_cfg2"] = {"model": "BASELINE configs[1]: COMBINATION, octree level 6", "wall_s": w2,
                                            "wall_with_host_reads_s": w2r, **info2}

    if rank == 0:
        n_o, n_rest, n_surf = st.n_ori, st.n_rest, st.n_surf
        F = flops_per_point(n_o, n_rest, n_surf)
        achieved_tf = value / world * F / 1e12                      # per GPU
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        nominal_peak = eng.sm_count * 64 * 2 * 1.965e9 / 1e12       # 148 SM x 64 FP64 lanes x 2 x 1.965 GHz
        peak_at_clock = eng.sm_count * 64 * 2 * sm_mhz * 1e6 / 1e12
        pcs = ClockSampler(local)                   # the probes' own clock record (they run for ~0.5 s in total)
        pcs.start()
        peaks = fp64_peaks(eng)
        peaks["clocks"] = pcs.stop()
        measured_peak = peaks["dfma_tflops"]
        # DRAM traffic of one launch of the dominant kernel: from the committed ncu --set full capture of this command
        # (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py), never a constant in this file
        traffic, traffic_src = None, None
        tf = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tf) and world == 1:
            rec = json.load(open(tf)).get(f"eval_zrun_grid{args.grid}")
            if rec:
                traffic, traffic_src = rec["dram_bytes_per_launch"], rec["source"]
        roof = {"bound": "fp64", "achieved": achieved_tf, "peak": measured_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / measured_peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": 32 * m,
                "peak_source": "measured in this run: DFMA-chain microbenchmark gpb_bench_dfma (MEASURED_PEAKS.json has no FP64 entry)",
                "peak_nominal": nominal_peak, "frac_of_nominal": achieved_tf / nominal_peak,
                "peak_at_measured_clock": peak_at_clock, "frac_at_measured_clock": achieved_tf / peak_at_clock,
                "flops_per_point": F, "kernel": "eval_zrun_kernel<cubic, grad, P=8, T=256>",
                "bound_note": "neither HBM (0.06 % DRAM utilisation) nor tensor: the FP64 FMA pipe is the binding resource (ncu: 85 % pipe-active, profiles/r2_eval_zrun_kernel_ncu_full_512.txt)",
                "hbm": {"achieved_gbs": value / world * 32 / 1e9, "algorithmic_bytes_per_point": 32}}
        cpu = None
        if not args.no_cpu and world == 1:
            v, dt, cores = cpu_baseline(model, args.cpu_sample)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_sample} of {n_total} grid points ({dt:.1f} s), " + _cpu_desc("numpy", 500_000)}
            # the same port at a chunk size that is not Python-overhead bound, and the PyTorch-CPU variant (north_star)
            v2, dt2, _ = cpu_baseline(model, args.cpu_sample, chunk_elems=32_000_000)
            v3, dt3, _ = cpu_baseline(model, args.cpu_sample, backend="torch", chunk_elems=32_000_000)
            cpu["variants"] = [
                {"value": v2, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{args.cpu_sample} points ({dt2:.1f} s), " + _cpu_desc("numpy", 32_000_000)},
                {"value": v3, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{args.cpu_sample} points ({dt3:.1f} s), " + _cpu_desc("torch", 32_000_000)},
            ]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, model),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": None if e2e_val is None else {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                                    "d2h_bytes_per_step": d2h, "steps_timed": max(1, args.steps),
                                                    "step_s": {"median": float(np.median(step_s)), "mean": float(np.mean(step_s)), "max": float(np.max(step_s))},
                                                    "value_from": "median over the steps of the per-step max over ranks",
                                                    "call": "gempy_b200.engine.compute.compute_dense_fields (streaming form of compute_model + compute_scalar_gradient for outputs too large to keep)"},
                "roofline": roof, "fp64_peaks": peaks, "cpu_baseline": cpu}
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--backend", default="numpy", choices=["numpy", "torch"], help="reference arm: numpy or PyTorch-CPU port")
    ap.add_argument("--cpu-chunk", type=int, default=500_000, help="reference arm: kernel-matrix elements per chunk (evaluation_chunk_size)")
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--sp-per-surface", type=int, default=1000)
    ap.add_argument("--n-ori", type=int, default=1000)
    ap.add_argument("--cpu-sample", type=int, default=131072)
    ap.add_argument("--cfg4-levels", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
