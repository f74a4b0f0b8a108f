"""Item-axis sharding across the GPUs of one box.

Items [0, NI) are split into ``world`` contiguous ranges (contiguous so that
"lower global index wins ties" survives concatenating shards in rank order).
Each rank scores its shard for the same user block (``pxr_score_topk``), the
per-shard top-K lists (fp32 score, int32 global index: 8*K bytes per user) are
exchanged with ONE collective and merged (``pxr_merge_topk``).  Two forms:

  * all-gather (``allgather_topk*``): every rank ends up with the merged lists of
    the whole block (what the evaluators use: metric sums then need no list
    traffic);
  * owned exchange (``exchange_owned_*``, an all-to-all): rank r receives only
    the lists of ITS 1/world slice of the block's users and merges those --
    reduce-scatter ownership: 1/world of the traffic and
    of the merge work per rank, the merged block stays distributed over the ranks.

No other communication is on the data path.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous ceil(NI/world)-sized ranges; trailing ranks may be short or empty."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def allgather_topk(scores: torch.Tensor, idx: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(n, K) local lists -> (world, n, K) stacked in rank order."""
    world = dist.get_world_size(group)
    n, k = scores.shape
    # the concatenated (world*n, K) form is the one both NCCL and gloo accept
    all_s = torch.empty((world * n, k), dtype=scores.dtype, device=scores.device)
    all_i = torch.empty((world * n, k), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(all_s, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, idx.contiguous(), group=group)
    return all_s.view(world, n, k), all_i.view(world, n, k)


def allgather_topk_start(scores: torch.Tensor, idx: torch.Tensor, group=None):
    """Start ONE asynchronous all-gather of the local lists (score bits and indices interleaved in a single
    (n, K, 2) int32 buffer: 8*K bytes per user per rank).  Returns a handle for ``allgather_topk_finish``; the
    collective runs on the communicator's own stream and overlaps whatever is launched next."""
    world = dist.get_world_size(group)
    n, k = scores.shape
    local = torch.stack([scores.contiguous().view(torch.int32), idx.to(torch.int32)], dim=-1).contiguous()
    out = torch.empty((world * n, k, 2), dtype=torch.int32, device=scores.device)
    work = dist.all_gather_into_tensor(out, local, group=group, async_op=True)
    return work, out, (world, n, k), local


def allgather_topk_finish(handle) -> Tuple[torch.Tensor, torch.Tensor]:
    """Wait for a started all-gather: (world, n, K) fp32 scores and int32 indices stacked in rank order."""
    work, out, (world, n, k), _local = handle
    work.wait()
    out = out.view(world, n, k, 2)
    return out[..., 0].contiguous().view(torch.float32), out[..., 1].contiguous()


def owned_slice(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Users [lo, hi) of an n-user block that rank ``rank`` owns after ``exchange_owned_*``."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def exchange_owned_start(scores: torch.Tensor, idx: torch.Tensor, group=None):
    """Start ONE asynchronous all-to-all of the local per-shard lists: rank r receives, from every rank, the lists of
    the users it owns (``owned_slice``).  Same packed (score bits, index) int32 rows as ``allgather_topk_start``; the
    block is padded to a multiple of ``world`` users with empty lists."""
    world = dist.get_world_size(group)
    n, k = scores.shape
    per = (n + world - 1) // world
    local = torch.empty((world * per, k, 2), dtype=torch.int32, device=scores.device)
    local[:n, :, 0] = scores.contiguous().view(torch.int32)
    local[:n, :, 1] = idx.to(torch.int32)
    if world * per > n:
        local[n:, :, 0] = torch.tensor(float("-inf"), dtype=torch.float32).view(torch.int32).item()
        local[n:, :, 1] = -1
    out = torch.empty((world * per, k, 2), dtype=torch.int32, device=scores.device)
    work = dist.all_to_all_single(out, local, group=group, async_op=True)
    return work, out, (world, per, k, n), local


def exchange_owned_finish(handle, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Wait for a started owned exchange: (world, m, K) fp32 scores and int32 indices of this rank's m owned users,
    stacked in shard (rank) order, ready for ``pxr_merge_topk``."""
    work, out, (world, per, k, n), _local = handle
    work.wait()
    lo, hi = owned_slice(n, world, dist.get_rank(group))
    out = out.view(world, per, k, 2)[:, :hi - lo]
    return out[..., 0].contiguous().view(torch.float32), out[..., 1].contiguous()


def gather_owned(scores: torch.Tensor, idx: torch.Tensor, n: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Owned final lists of an n-user block ((m, K) per rank, ``owned_slice``) -> the (n, K) lists of the whole block on every
    rank (one small all-gather: K * 8 bytes per user)."""
    world = dist.get_world_size(group)
    per = (n + world - 1) // world
    k = scores.shape[1]
    local = torch.empty((per, k, 2), dtype=torch.int32, device=scores.device)
    m = scores.shape[0]
    local[:m, :, 0] = scores.contiguous().view(torch.int32)
    local[:m, :, 1] = idx.to(torch.int32)
    if m < per:
        local[m:] = 0
    out = torch.empty((world * per, k, 2), dtype=torch.int32, device=scores.device)
    dist.all_gather_into_tensor(out, local, group=group)
    out = out[:n]
    return out[..., 0].contiguous().view(torch.float32), out[..., 1].contiguous()


class ShardedTopK:
    """Lock-step sharded scoring: ``local_topk(users, k, filter_seen)`` is the rank's scorer over its item range
    (``FastRecommender.recommend_all`` in production), ``merge`` the S-way merge (``pxr_merge_topk``).

    Exact mode across shards (``rescore`` given, e.g. ``FastRecommender.rescore``): ``local_topk`` must then return the RAW
    16-bit lists (``recommend_all(..., raw=True)``); the exchange carries their 64 slots, the rank that owns a user merges
    the shards' lists into the global 16-bit top-64 and re-scores THOSE in fp32 against whole-catalogue records --
    the same candidates, hence the same lists, as the unsharded exact mode, and 1 / world of the re-score work per rank
    instead of every shard re-scoring every user.  ``top_k`` > 64: the raw lists have 64 * ceil(top_k / 64) slots (one
    pass of the fused kernel per 64-slot page), exchanged, merged and re-scored the same way."""

    RAW_K = 64

    def __init__(self, local_topk: Callable, merge: Optional[Callable] = None, group=None, rescore: Optional[Callable] = None):
        self.local_topk = local_topk
        if merge is None:
            from .engine import merge_topk as merge
        self.merge = merge
        self.group = group
        self.rescore = rescore

    def _sharded(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _k_exchange(self, top_k: int) -> int:
        return self.RAW_K * ((top_k + self.RAW_K - 1) // self.RAW_K) if self.rescore is not None else top_k

    def _finish_owned(self, handle, blk, top_k: int):
        s, i = self.merge(*exchange_owned_finish(handle, self.group))
        if self.rescore is not None:
            lo, hi = owned_slice(len(blk), dist.get_world_size(self.group), dist.get_rank(self.group))
            s, i = self.rescore(blk[lo:hi], i, top_k)
        return s, i

    def recommend_blocks_owned(self, user_blocks, top_k: int, filter_seen: bool = True):
        """Generator over user blocks with reduce-scatter ownership: yields, per block, the final lists of the users THIS
        rank owns (``owned_slice(len(block), world, rank)``).  The exchange of block b (one all-to-all on the
        communicator's stream) overlaps the scoring of block b + 1: the local kernel of the next block is launched before
        the previous block's exchange is waited for, merged and re-scored."""
        pending = None
        for blk in user_blocks:
            s, i = self.local_topk(blk, self._k_exchange(top_k), filter_seen)
            if not self._sharded():
                if self.rescore is not None:
                    s, i = self.rescore(blk, i, top_k)
                yield s, i
                continue
            handle = exchange_owned_start(s, i, self.group)
            if pending is not None:
                yield self._finish_owned(pending[0], pending[1], top_k)
            pending = (handle, blk)
        if pending is not None:
            yield self._finish_owned(pending[0], pending[1], top_k)

    def recommend_blocks(self, user_blocks, top_k: int, filter_seen: bool = True):
        """As ``recommend_blocks_owned``, then the owned lists are gathered so that every rank holds the whole block."""
        user_blocks = list(user_blocks)
        for blk, (s, i) in zip(user_blocks, self.recommend_blocks_owned(user_blocks, top_k, filter_seen)):
            if self._sharded():
                s, i = gather_owned(s, i, len(blk), self.group)
            yield s, i

    def recommend_all(self, user_indices, top_k: int, filter_seen: bool = True):
        return next(iter(self.recommend_blocks([user_indices], top_k, filter_seen)))
