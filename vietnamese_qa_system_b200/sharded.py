"""Row-sharded exact search across the GPUs of one box.

Rank r of G holds the contiguous row block ``[r*ceil(N/G), min((r+1)*ceil(N/G), N))``
of the document matrix.  A search is: local scan + fused top-k on every rank ->
ONE all-gather of the packed ``[B, k]`` candidates (NCCL over NVLink/NVSwitch,
through ``torch.distributed``) -> merge-top-k kernel (K4) -> identical ``[B, k]`` on
every rank.  There is no other collective on the data path.

The reference is single-process (heavy_ranker.py:97-101); this module is what
lets its one index grow past one GPU.  The partition arithmetic and the
exchange are backend-agnostic so that world_size-2 ``gloo`` tests can drive them
on CPU with the oracle standing in for the two device steps (tests only).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of ``rank``: [lo, hi)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    per = -(-n_total // world) if n_total > 0 else 0
    lo = min(rank * per, n_total)
    hi = min(lo + per, n_total)
    return lo, hi


def exchange_candidates(scores: torch.Tensor, ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One all-gather of the per-rank ``[B, k]`` results.

    Scores (fp32) and ids (int64) are packed into a single int64 ``[B, k, 2]`` buffer so
    that exactly one collective is issued.  Returns ``([G,B,k] scores, [G,B,k] ids)``.
    """
    world = dist.get_world_size(group)
    b, k = scores.shape
    packed = torch.empty((b, k, 2), dtype=torch.int64, device=scores.device)
    packed[..., 0] = scores.contiguous().view(torch.int32).to(torch.int64)
    packed[..., 1] = ids
    flat = torch.empty((world * b, k, 2), dtype=torch.int64, device=scores.device)
    dist.all_gather_into_tensor(flat, packed, group=group)  # rank-major concatenation along dim 0
    gathered = flat.view(world, b, k, 2)
    g_scores = gathered[..., 0].to(torch.int32).view(torch.float32).contiguous()
    g_ids = gathered[..., 1].contiguous()
    return g_scores, g_ids


class ShardedSearch:
    """Host-side driver of the sharded search; device steps are injected.

    ``local_search(queries, k) -> (scores [B,k], ids [B,k])`` with GLOBAL ids, padded with
    ``(-inf, -1)``; ``merge(cand_scores [G,B,k], cand_ids [G,B,k], k) -> (scores, ids)``.
    """

    def __init__(self, local_search: Callable, merge: Callable, group=None):
        self.local_search = local_search
        self.merge = merge
        self.group = group

    def search(self, queries: torch.Tensor, k: int):
        s, i = self.local_search(queries, k)
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return s, i
        gs, gi = exchange_candidates(s, i, self.group)
        return self.merge(gs, gi, k)


class ShardedFlat:
    """The product wiring: ``ops.FlatShard`` scan + one exchange of the packed ``[B,k]`` results + merge (K4).

    One process per GPU (torchrun); ``rows`` is THIS rank's block, already in storage dtype.
    The local search writes its ``[B,k]`` scores and ids straight into one packed byte block; the merge
    kernel reads the gathered blocks in place (no repacking).  Two interchangeable exchanges:

    * ``exchange="nccl"``: ONE ``all_gather_into_tensor`` (NCCL over NVLink/NVSwitch) -- 4 kernels per search;
    * ``exchange="p2p"`` : a push kernel stores the block into every peer's gather buffer through NVLink
      peer memory (torch symmetric memory) and publishes a release flag; the merge kernel acquires the
      flags -- no NCCL launch on the data path.  ``exchange="auto"`` tries p2p and falls back to NCCL
      when symmetric memory is not available.  Default: ``"nccl"`` (BASELINE.json north_star's design;
      measured within 1-2 % of p2p at the sizes of interest).
    """

    def __init__(self, rows: torch.Tensor, n_total: int, group=None, mode="fast", exchange: str = "nccl"):
        from . import ops

        self._ops = ops
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        if rows.shape[0] != hi - lo:
            raise ValueError(f"rank {self.rank} must hold rows [{lo},{hi}) = {hi - lo} rows; got {rows.shape[0]}")
        self.n_total = n_total
        self.shard = ops.FlatShard(rows, first_global_id=lo)
        self.mode = mode
        if exchange not in ("auto", "nccl", "p2p"):
            raise ValueError(f"exchange must be auto, nccl or p2p; got {exchange!r}")
        self.exchange = exchange
        self._bufs = {}
        self._epoch = 0

    @staticmethod
    def _layout(b: int, k: int):
        ids_off = (b * k * 4 + 15) // 16 * 16
        block = (ids_off + b * k * 8 + 15) // 16 * 16
        return ids_off, block

    def _buffers(self, b: int, k: int):
        key = (b, k)
        buf = self._bufs.get(key)
        if buf is None:
            dev = self.shard.device
            ids_off, block = self._layout(b, k)
            local = torch.zeros(block, dtype=torch.uint8, device=dev)
            s_view = local[:b * k * 4].view(torch.float32).view(b, k)
            i_view = local[ids_off:ids_off + b * k * 8].view(torch.int64).view(b, k)
            out_s = torch.empty((b, k), dtype=torch.float32, device=dev)
            out_i = torch.empty((b, k), dtype=torch.int64, device=dev)
            p2p = None
            if self.exchange in ("auto", "p2p") and k <= 32 and self.world <= 16:
                try:
                    p2p = self._ops.PeerExchange(block, self.rank, self.world, dev, self.group)
                except Exception:  # noqa: BLE001 - symmetric memory unavailable on this build / topology
                    if self.exchange == "p2p":
                        raise
                    p2p = None
            gathered = None if p2p is not None else torch.empty(block * self.world, dtype=torch.uint8, device=dev)
            buf = (local, gathered, s_view, i_view, ids_off, out_s, out_i, p2p)
            self._bufs[key] = buf
        return buf

    def search(self, queries: torch.Tensor, k: int, mode: Optional[str] = None):
        """Returns (scores [B,k], ids [B,k]) -- identical on every rank.  The returned tensors are
        reused by the next search with the same (B, k)."""
        if mode is not None:
            self.mode = mode
        if self.world == 1:
            return self.shard.search(queries, k, self.mode)
        b = int(queries.shape[0]) if queries.dim() == 2 else 1
        local, gathered, s_view, i_view, ids_off, out_s, out_i, p2p = self._buffers(b, k)
        self.shard.search(queries, k, self.mode, s_view, i_view)
        if p2p is not None:
            return p2p.push_and_merge(local, b, k, ids_off, out_s, out_i)
        dist.all_gather_into_tensor(gathered, local, group=self.group)
        return self._ops.merge_topk_packed(gathered, self.world, b, k, ids_off, out_s, out_i)
