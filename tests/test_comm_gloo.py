"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: sharding, weight broadcast, fault-minimum all-reduce,
refine-mark all-gather.  The NCCL path on real GPUs is exercised by tests/dist_compute_check.py under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gempy_b200.engine.comm import Comm, shard_range, shard_sizes


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 1000, 134217728):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = Comm()
        assert comm.world == world and comm.rank == rank
        # refine marks: every rank tests its voxel range, the gather restores the global order
        nv = 1001
        marks_global = (torch.arange(nv) % 3 == 0).to(torch.uint8)
        v0, v1 = comm.shard(nv)
        full = comm.all_gather_cat(marks_global[v0:v1].clone(), nv)
        assert torch.equal(full, marks_global)
        # fields: 2-D tensors sharded along the last dimension (ragged: 7 columns over 2 ranks)
        ref = torch.arange(3 * 7, dtype=torch.float64).reshape(3, 7)
        c0, c1 = comm.shard(7)
        got = comm.all_gather_cat(ref[:, c0:c1].contiguous(), 7)
        assert torch.equal(got, ref)
        # level outputs gathered in place (CPU tensors take the all_gather_cat + copy route; NCCL rows land directly)
        ref3 = torch.arange(2 * 3 * 8, dtype=torch.float64).reshape(2, 3, 8)
        out3 = torch.zeros(2, 3, 11, dtype=torch.float64)
        c0, c1 = comm.shard(8)
        comm.all_gather_rows_into(out3[..., 2:10], ref3[..., c0:c1], 8)
        assert torch.equal(out3[..., 2:10], ref3) and out3[..., :2].abs().sum() == 0 and out3[..., 10:].abs().sum() == 0
        # an empty shard on one rank
        e0, e1 = comm.shard(1)
        got = comm.all_gather_cat(torch.full((e1 - e0,), 5.0), 1)
        assert got.tolist() == [5.0]
        # weights: rank 0 "solves", everybody ends up with its vector
        w = torch.arange(10, dtype=torch.float64) if rank == 0 else torch.zeros(10, dtype=torch.float64)
        comm.broadcast(w, src=0)
        assert torch.equal(w, torch.arange(10, dtype=torch.float64))
        # fault-block minimum over point shards
        m = torch.tensor([float(3 - rank)])
        comm.all_reduce_min(m)
        assert m.item() == 3.0 - (world - 1)
        # a level too small to shard: every rank keeps the whole range and nothing is exchanged
        solo = Comm.solo()
        assert (solo.world, solo.rank, solo.enabled) == (1, 0, False) and solo.shard(nv) == (0, nv)
        assert solo.all_gather_cat(marks_global, nv) is marks_global
        m = torch.tensor([float(rank)])
        assert solo.all_reduce_min(m).item() == float(rank)
        q.put((rank, "ok"))
    except Exception as exc:          # pragma: no cover
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_comm_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
