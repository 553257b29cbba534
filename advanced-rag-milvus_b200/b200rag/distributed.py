"""Row-sharded search across the GPUs of one box (one process per GPU, torch.distributed).

The corpus is cut into contiguous row ranges, one per rank; every rank scans its shard for
the exact local top-k, ONE all-gather carries the k (score, id) candidates per query and rank, and the merge kernel
(K4) reduces G*k -> k by (score desc, id asc) on every rank.  Scores travel as fp64 so the merge compares the same
canonical values the single-GPU path ranks by -- the merged result is bit-identical to a single-shard search.

The exchange is the only collective on the path (NCCL over NVLink/NVSwitch on the GPU box; the same code runs over
gloo in the CPU tests).  It is latency bound: B=1024, k=100 is 1.6 MB per rank.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced row range [start, end) of `rank`; the first n_total % world ranks hold one extra row."""
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_candidates(scores: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """(f64 [B,k], i64 [B,k]) -> one i64 [B, 2k] message (score bit patterns, then ids)."""
    return torch.cat([scores.contiguous().view(torch.int64), ids], dim=1).contiguous()


def unpack_gathered(gathered: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """i64 [G, B, 2k] -> candidate lists (f64 [B, G*k], i64 [B, G*k]), rank-major inside a query."""
    g, b, _ = gathered.shape
    sc = gathered[:, :, :k].permute(1, 0, 2).reshape(b, g * k).contiguous().view(torch.float64)
    ids = gathered[:, :, k:].permute(1, 0, 2).reshape(b, g * k).contiguous()
    return sc, ids


def gather_and_merge(scores: torch.Tensor, ids: torch.Tensor, k: int,
                     merge_fn: Callable[[torch.Tensor, torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]],
                     group: Optional[dist.ProcessGroup] = None,
                     merge_gathered_fn: Optional[Callable[[torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]]] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather every rank's local top-k and merge.  Identity when not running distributed.

    merge_gathered_fn (engine.merge_gathered on the GPU) reduces the gathered [G, B, 2k] buffer in place; without it the
    buffer is unpacked into [B, G*k] candidate lists for merge_fn (the CPU tests stand the oracle in there)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return scores, ids
    msg = pack_candidates(scores, ids)
    # concatenated (not stacked) output layout: the one form both NCCL and gloo accept
    out = torch.empty((world * msg.shape[0], msg.shape[1]), dtype=msg.dtype, device=msg.device)
    dist.all_gather_into_tensor(out, msg, group=group)
    gathered = out.view(world, msg.shape[0], msg.shape[1])
    if merge_gathered_fn is not None:
        return merge_gathered_fn(gathered, k)
    cs, ci = unpack_gathered(gathered, k)
    return merge_fn(cs, ci, k)


class ShardedDenseIndex:
    """DenseIndex over this rank's row range of a corpus of `n_total` rows."""

    def __init__(self, dim: int, n_total: int, dtype="f16", metric: str = "COSINE", device="cuda",
                 group: Optional[dist.ProcessGroup] = None):
        from . import engine
        self._engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_total = n_total
        self.start, self.end = shard_range(n_total, self.rank, self.world)
        self.local = engine.DenseIndex(dim, dtype, metric, device, id_offset=self.start, capacity=self.end - self.start)

    def search(self, queries_f32: torch.Tensor, k: int, mode: Optional[int] = None):
        """Replicated queries -> global exact top-k on every rank: (scores f64 [B,k], ids i64 [B,k])."""
        eng = self._engine
        s, i, _ = self.local.search(queries_f32, k, eng.DENSE_AUTO if mode is None else mode)
        return gather_and_merge(s, i, k, eng.merge_topk, self.group, eng.merge_gathered)
