#!/usr/bin/env python
"""compute_model wall time of the multi-fault model (BASELINE configs[3]: 10 fault stacks + 5 series, octree level 8,
dual contouring) with the octree levels sharded over the ranks.  Run under torchrun; rank 0 prints one JSON line."""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import examples as ex
from gempy_b200.engine import compute as gc
from gempy_b200.engine.comm import Comm

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=8)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = Comm()
eng = gc.B200Engine(local)
build = lambda: ex.synthetic_multi_fault(refinement=args.levels)
sol = gc.compute_model(*build().args(), engine=eng, comm=comm)      # warm-up
ts = []
for _ in range(args.reps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sol = gc.compute_model(*build().args(), engine=eng, comm=comm)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
t = torch.tensor([min(ts)], dtype=torch.float64, device=eng.device)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if comm.rank == 0:
    leaves = [int(l.grid_centers.octree_grid.n_points) for l in sol.octrees_output]
    ids = sol.octrees_output[-1].outputs_centers[-1].ids_block
    print(json.dumps({"model": f"multi_fault_10f_5s_octree{args.levels}", "n_gpus": world, "compute_model_wall_s": float(t.item()),
                      "leaf_counts": leaves, "n_meshes": len(sol.dc_meshes), "ids_checksum": float(np.asarray(ids).sum())}))
if world > 1:
    dist.destroy_process_group()
