"""Multi-GPU plumbing of the B200 backend: one process per GPU, torch.distributed (NCCL over NVLink on the
GPUs; gloo in the CPU tests).

The path shards by evaluation-point range: every rank evaluates all stacks on its contiguous
slice of the regular grid / of the octree level's voxel list.  The only exchanges are
  * nothing for the weights: every rank assembles and solves every stack itself (the symmetric solve takes 8 ms at
    n = 7 000 -- cheaper than shipping weights around -- and identical inputs give bit-identical weights),
  * all-reduce(MIN) of each fault block's minimum (one double per fault stack and level),
  * all-gather of the per-level refine marks (1 byte per voxel), so that every rank builds the identical child
    list -- the leaf order, and therefore every output, is independent of the number of GPUs,
  * all-gather of the output slices when the host asks for whole arrays.
Levels that do not pay for these exchanges (compute._shard_pays: evaluation time saved against the all-gather of the level's
outputs) are not sharded: every rank evaluates them whole (`Comm.solo()`), which leaves the ranks in the identical state.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra element."""
    base, rem = divmod(int(n), int(world))
    i0 = rank * base + min(rank, rem)
    return i0, i0 + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


class Comm:
    """No-op when torch.distributed is not initialised (single GPU)."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world = dist.get_world_size(group) if self.enabled else 1

    @classmethod
    def solo(cls) -> "Comm":
        """A Comm of one rank whatever the process group: used for work too small to shard (every rank repeats it and
        holds the identical result, nothing is exchanged)."""
        c = cls.__new__(cls)
        c.group, c.enabled, c.rank, c.world = None, False, 0, 1
        return c

    def shard(self, n: int) -> Tuple[int, int]:
        return shard_range(n, self.rank, self.world)

    def broadcast(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.enabled and self.world > 1:
            dist.broadcast(t, src=src, group=self.group)
        return t

    def all_reduce_min(self, t: torch.Tensor) -> torch.Tensor:
        if self.enabled and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.enabled and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_gather_rows_into(self, out: torch.Tensor, local: torch.Tensor, n_total: int) -> None:
        """out[..., :n_total] = the ranks' shards of `local` (range(n_total)-sharded along the LAST dimension) in rank
        order.  With equal shards on CUDA tensors every leading-index row is gathered by NCCL straight into its place
        (no padded staging copy, no concatenation); otherwise all_gather_cat + one copy."""
        if not (self.enabled and self.world > 1):
            out[..., :n_total].copy_(local)
            return
        if n_total == 0:
            return
        if n_total % self.world == 0 and local.is_cuda and hasattr(dist, "all_gather_into_tensor"):
            m = n_total // self.world
            assert local.shape[-1] == m, (local.shape, n_total, self.world)
            try:                                  # views only: a reshape that copies would swallow the gathered rows
                o2, l2 = out.view(-1, out.shape[-1]), local.view(-1, local.shape[-1])
            except RuntimeError:
                o2 = l2 = None
            if o2 is not None and o2.stride(-1) == 1 and l2.stride(-1) == 1 and o2.shape[0] == l2.shape[0]:
                for r in range(o2.shape[0]):
                    dist.all_gather_into_tensor(o2[r, :n_total], l2[r], group=self.group)
                return
        out[..., :n_total].copy_(self.all_gather_cat(local.contiguous(), n_total))

    def all_gather_cat(self, local: torch.Tensor, n_total: int) -> torch.Tensor:
        """Concatenate the ranks' shards of a range(n_total)-sharded tensor along its LAST dimension.
        Shard sizes follow shard_range, so no size exchange is needed."""
        if not (self.enabled and self.world > 1):
            return local
        sizes = shard_sizes(n_total, self.world)
        assert local.shape[-1] == sizes[self.rank], (local.shape, sizes, self.rank)
        mx = max(sizes)
        lead = local.shape[:-1]
        padded = torch.zeros(*lead, mx, dtype=local.dtype, device=local.device)
        padded[..., :local.shape[-1]] = local
        out = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(out, padded.contiguous(), group=self.group)
        return torch.cat([o[..., :s] for o, s in zip(out, sizes)], dim=-1)
