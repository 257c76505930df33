"""Run under torchrun on N GPUs: compute_model sharded over the ranks must return, on every rank, exactly what a
single-GPU run returns (leaf order, ids, fields).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_compute_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gempy_b200 import examples as ex                    # noqa: E402
from gempy_b200.engine import compute as gc              # noqa: E402
from gempy_b200.engine.comm import Comm                  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm()
    eng = gc.B200Engine(local)
    worst = 0.0
    builds = (ex.combination, ex.one_fault, lambda: ex.combination(resolution=(21, 10, 9)), lambda: ex.combination(resolution=(7, 5, 3)),
              lambda: ex.greenstone(refinement=4))          # (7, 5, 3): 105 dense points, ragged shards
    # every level sharded ("0"), only the deeper levels sharded, and the default threshold (these models: nothing sharded)
    for build, min_pairs in [(b, t) for b in builds for t in ("0", "2e5", None)]:
        if min_pairs is None:
            os.environ.pop("GPB_SHARD_MIN_PAIRS", None)
        else:
            os.environ["GPB_SHARD_MIN_PAIRS"] = min_pairs
        sol_d = gc.compute_model(*build().args(), engine=eng, comm=comm)
        sol_1 = gc.compute_model(*build().args(), engine=eng, comm=None if False else _Single())
        assert len(sol_d.octrees_output) == len(sol_1.octrees_output)
        for a, b in zip(sol_d.octrees_output, sol_1.octrees_output):
            np.testing.assert_array_equal(a.grid_centers.octree_grid.values, b.grid_centers.octree_grid.values)
            for oa, ob in zip(a.outputs_centers, b.outputs_centers):
                za, zb = oa.exported_fields.scalar_field_everywhere, ob.exported_fields.scalar_field_everywhere
                assert za.shape == zb.shape
                worst = max(worst, float(np.abs(za - zb).max()))
                np.testing.assert_array_equal(oa.scalar_fields.values_block, ob.scalar_fields.values_block)
            np.testing.assert_array_equal(a.outputs_centers[-1].block, b.outputs_centers[-1].block)
        np.testing.assert_array_equal(sol_d.raw_arrays.lith_block, sol_1.raw_arrays.lith_block)
        if sol_1.dc_meshes is not None:
            for ma, mb in zip(sol_d.dc_meshes, sol_1.dc_meshes):
                np.testing.assert_allclose(ma.vertices, mb.vertices, rtol=0, atol=1e-12)
                np.testing.assert_array_equal(ma.edges, mb.edges)
    t = torch.tensor([worst], device=eng.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if comm.rank == 0:
        print(f"dist_compute_check ok on {comm.world} GPUs: fields identical to the single-GPU run (max |dZ| = {t.item():.1e})")
    dist.destroy_process_group()


class _Single(Comm):
    """A Comm that ignores the process group: every rank computes the whole model alone."""
    def __init__(self):
        self.group, self.enabled, self.rank, self.world = None, False, 0, 1


if __name__ == "__main__":
    main()
