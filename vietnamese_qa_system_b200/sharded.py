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
        self._side = None      # side stream of the pipelined search (exchange + merge of batch i under scan i+1)
        self._copy = None      # copy stream of the host-buffer form (H2D of batch i+1 under scan i)
        self._last = {}        # (b, k, slot) -> event after which the slot's buffers may be reused

    @staticmethod
    def _layout(b: int, k: int):
        ids_off = (b * k * 4 + 15) // 16 * 16
        block = (ids_off + b * k * 8 + 15) // 16 * 16
        return ids_off, block

    def _buffers(self, b: int, k: int, slot: int = 0):
        """Per (B, k, slot): the packed local result block, the gather buffer (or the peer-memory exchange), the
        merged outputs and a PRIVATE scan workspace -- two slots never share scratch, so searches of different
        slots may be in flight together (and searches issued from different host threads must use different slots)."""
        key = (b, k, slot)
        buf = self._bufs.get(key)
        if buf is None:
            dev = self.shard.device
            ids_off, block = self._layout(b, k)
            local = torch.zeros(block, dtype=torch.uint8, device=dev)
            s_view = local[:b * k * 4].view(torch.float32).view(b, k)
            i_view = local[ids_off:ids_off + b * k * 8].view(torch.int64).view(b, k)
            out_s = torch.empty((b, k), dtype=torch.float32, device=dev)
            out_i = torch.empty((b, k), dtype=torch.int64, device=dev)
            ws = torch.zeros(self.shard.workspace_bytes(b, k, self.mode), dtype=torch.uint8, device=dev)
            p2p = None
            if self.exchange in ("auto", "p2p") and self.world <= 16:
                try:
                    p2p = self._ops.PeerExchange(block, self.rank, self.world, dev, self.group)
                except Exception:  # noqa: BLE001 - symmetric memory unavailable on this build / topology
                    if self.exchange == "p2p":
                        raise
                    p2p = None
            gathered = None if p2p is not None else torch.empty(block * self.world, dtype=torch.uint8, device=dev)
            buf = (local, gathered, s_view, i_view, ids_off, out_s, out_i, p2p, ws)
            self._bufs[key] = buf
        return buf

    def _exchange(self, buf, b: int, k: int):
        local, gathered, _, _, ids_off, out_s, out_i, p2p, _ = buf
        if p2p is not None:
            return p2p.push_and_merge(local, b, k, ids_off, out_s, out_i)
        dist.all_gather_into_tensor(gathered, local, group=self.group)
        return self._ops.merge_topk_packed(gathered, self.world, b, k, ids_off, out_s, out_i)

    def search(self, queries: torch.Tensor, k: int, mode: Optional[str] = None):
        """Returns (scores [B,k], ids [B,k]) -- identical on every rank, valid in stream order on the current
        stream.  The returned tensors are reused by the next search with the same (B, k)."""
        if mode is not None:
            self.mode = mode
        if self.world == 1:
            return self.shard.search(queries, k, self.mode)
        b = int(queries.shape[0]) if queries.dim() == 2 else 1
        if k > self._ops.K_CALL_MAX:
            # k > 128 (hybrid search with limit > 12): every rank composes its shard's top k from row segments
            # (ops.FlatShard._search_wide), the lists are all-gathered and merged by the same kernel -- ranks are
            # row ranges in ascending order, exactly the segments of vqa_merge_segments
            s, i = self.shard.search(queries, k, self.mode)
            gs = torch.empty((self.world, b, k), dtype=torch.float32, device=s.device)
            gi = torch.empty((self.world, b, k), dtype=torch.int64, device=s.device)
            dist.all_gather_into_tensor(gs, s.contiguous(), group=self.group)
            dist.all_gather_into_tensor(gi, i.contiguous(), group=self.group)
            out_s, out_i, _ = self._ops.merge_segments(gs, gi, k)
            return out_s, out_i
        buf = self._buffers(b, k)
        self.shard.search(queries, k, self.mode, buf[2], buf[3], workspace=buf[8])
        return self._exchange(buf, b, k)

    def search_pipelined(self, queries: torch.Tensor, k: int, slot: int):
        """One search whose exchange + merge run on a side stream, so that the NEXT call's scan (another ``slot``)
        overlaps them: the scan of batch i+1 streams the shard while batch i's 384-byte-per-query candidates cross
        NVLink and are merged.  Returns ``(scores, ids, event)``; the tensors hold the result once ``event`` has
        completed (``event.synchronize()`` on the host, or ``stream.wait_event(event)``), and belong to ``slot``
        until the next call with the same slot.  Alternate slots 0 / 1.  ``queries`` must stay untouched until the
        scan has run (stream order on the current stream)."""
        b = int(queries.shape[0]) if queries.dim() == 2 else 1
        dev = self.shard.device
        main = torch.cuda.current_stream(dev)
        if self.world == 1:
            ws = self._bufs.get(("ws1", b, k, slot))
            if ws is None:
                ws = (torch.zeros(self.shard.workspace_bytes(b, k, self.mode), dtype=torch.uint8, device=dev),
                      torch.empty((b, k), dtype=torch.float32, device=dev),
                      torch.empty((b, k), dtype=torch.int64, device=dev))
                self._bufs[("ws1", b, k, slot)] = ws
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            prev = self._last.get((b, k, slot))
            if prev is not None:
                main.wait_event(prev)      # the slot's previous reduce has read its candidate lists
            # scan on the current stream, candidate reduce on the side stream: the next call's scan follows this
            # one back to back (vqa_search_2s)
            s, i = self.shard.search(queries, k, self.mode, ws[1], ws[2], workspace=ws[0], reduce_stream=self._side)
            ev = torch.cuda.Event()
            ev.record(self._side)
            self._last[(b, k, slot)] = ev
            return s, i, ev
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        buf = self._buffers(b, k, slot)
        if buf[7] is None:
            # NCCL exchange: its kernel cannot share an SM with the scan's 226 KB CTAs and spins until every rank has
            # joined, so on a side stream it would sit on an SM the next scan needs (measured on 8 GPUs, config D:
            # 6.64 ms pipelined vs 5.91 ms in order) -- keep the step in stream order
            self.shard.search(queries, k, self.mode, buf[2], buf[3], workspace=buf[8])
            out_s, out_i = self._exchange(buf, b, k)
            done = torch.cuda.Event()
            done.record(main)
            self._last[(b, k, slot)] = done
            return out_s, out_i, done
        prev = self._last.get((b, k, slot))
        if prev is not None:
            main.wait_event(prev)          # the slot's previous exchange has read `local` and written the outputs
        # scan on the current stream; candidate reduce, push and merge on the side stream (vqa_search_2s hands over
        # with an event recorded right after the scan), so consecutive scans run back to back
        self.shard.search(queries, k, self.mode, buf[2], buf[3], workspace=buf[8], reduce_stream=self._side)
        with torch.cuda.stream(self._side):
            out_s, out_i = self._exchange(buf, b, k)
            done = torch.cuda.Event()
            done.record(self._side)
        self._last[(b, k, slot)] = done
        return out_s, out_i, done

    def search_host_pipelined(self, queries_host: torch.Tensor, k: int, slot: int, two_stream: Optional[bool] = None):
        """``search_pipelined`` with HOST buffers: pinned float32 ``[B, dim]`` queries in, pinned ``(scores, ids)`` out.
        Enqueues the H2D copy of the queries, the scan, the exchange + merge and the D2H copy of the merged result and
        returns ``(scores_host, ids_host, event)`` at once; the host tensors hold the result when ``event`` has
        completed.  A serving loop keeps two or three slots in flight (three measured best at 0.29 ms steps, bench.py
        ``--inflight``): it reads slot s's result (``event.synchronize()``)
        just before re-issuing slot s, so the per-step host wake-up is off the critical path while every byte of
        every step still crosses PCIe inside the loop.  ``queries_host`` must be pinned and stay untouched until
        ``event`` has completed."""
        if queries_host.is_cuda or queries_host.dtype != torch.float32 or not queries_host.is_contiguous():
            raise ValueError("queries_host must be a contiguous float32 CPU tensor")
        if not queries_host.is_pinned():
            raise ValueError("queries_host must be pinned (page-locked) for an asynchronous copy")
        if not (self.world > 1 if two_stream is None else two_stream):
            # one GPU, one call: vqa_search_host_async (H2D, scan, reduce, D2H in stream order on the current stream).
            # two_stream=True takes the path below on one GPU as well: copy stream + vqa_search_2s + D2H behind the reduce
            return self.shard.search_host_async(queries_host, k, self.mode, slot=slot)
        b = int(queries_host.shape[0])
        dev = self.shard.device
        key = ("host", b, k, slot)
        st = self._bufs.get(key)
        if st is None:
            st = (torch.empty((b, self.shard.dim), dtype=torch.float32, device=dev),
                  torch.empty((b, k), dtype=torch.float32).pin_memory(),
                  torch.empty((b, k), dtype=torch.int64).pin_memory())
            self._bufs[key] = st
        q_dev, hs, hi = st
        main = torch.cuda.current_stream(dev)
        if self._copy is None:
            self._copy = torch.cuda.Stream(device=dev)
        prev = self._last.get((b, k, slot))
        with torch.cuda.stream(self._copy):
            # the queries of step i+1 cross PCIe on a copy stream while the scan of step i runs; the scan only waits
            # for the event (the slot's previous D2H copies are done before q_dev is re-written)
            if prev is not None:
                self._copy.wait_event(prev)
            q_dev.copy_(queries_host, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self._copy)
        main.wait_event(copied)
        out_s, out_i, done = self.search_pipelined(q_dev, k, slot)
        with torch.cuda.stream(self._side):
            self._side.wait_event(done)    # (the NCCL form finishes on the main stream)
            hs.copy_(out_s, non_blocking=True)
            hi.copy_(out_i, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._side)
        self._last[(b, k, slot)] = done
        return hs, hi, done
