"""ANN-level plugin: ``B200Flat`` -- exact (flat, inner-product) search on one B200.

Mirrors the txtai ANN backend interface the reference's retriever sits on
(``index / append / delete / search / count / save / load``; the reference reaches
it through ``txtai.Embeddings`` at heavy_ranker.py:78-101, where txtai would
otherwise build a faiss ``IDMap,Flat`` / ``IVFx,Flat`` CPU index).  A txtai
install can mount it with ``backend="vietnamese_qa_system_b200.ann.B200Flat"``.

Row position is the ANN id (txtai: ``np.arange(N)``); positions are stable across
``delete`` (faiss ``IDMap.remove_ids`` semantics).  All arithmetic is in
``libvqa_b200.so``; there is no CPU path.
"""
from __future__ import annotations

import json
import os
import warnings
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

_FORMAT_VERSION = 1


def _as_cuda_f32(x, device: torch.device) -> torch.Tensor:
    """Host/device array-like -> float32 CUDA tensor (a copy engine job, no arithmetic)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError(f"expected a [n, dim] array; got shape {tuple(t.shape)}")
    return t.to(device=device, dtype=torch.float32, non_blocking=True)


class B200Flat:
    """Exact inner-product index resident in one GPU's HBM."""

    def __init__(self, config: Optional[dict] = None):
        self.config = dict(config or {})
        dt = str(self.config.get("dtype", "bf16")).lower()
        if dt not in ops.DTYPES:
            raise ValueError(f"dtype must be one of {sorted(ops.DTYPES)}; got {dt!r}")
        self.dtype = ops.DTYPES[dt]
        self.mode = self.config.get("mode", "verify" if self.dtype == torch.float32 else "fast")
        dev = self.config.get("device", None)
        self._device_arg = dev
        self.normalize = bool(self.config.get("normalize", False))  # vectors arrive normalised (txtai does it upstream)
        self.first_global_id = int(self.config.get("first_global_id", 0))
        self.shard: Optional[ops.FlatShard] = None
        self._positions: Optional[torch.Tensor] = None  # row -> ANN id once rows were deleted
        self._next_id = 0
        # capacity-reserved storage: `_buf` holds `_cap` rows of which the first `shard.n` are live, so that
        # append() writes new rows behind the live ones in place (amortised O(rows appended), no copy of the index)
        self._buf: Optional[torch.Tensor] = None
        self.growth = float(self.config.get("growth", 1.5))   # capacity factor when the buffer has to grow
        self.compact_chunk = int(self.config.get("compact_chunk", 1 << 20))  # rows moved per step by delete()

    # ------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        from . import _native

        _native.require_cuda()
        d = self._device_arg
        if d is None:
            return torch.device("cuda", torch.cuda.current_device())
        d = torch.device(d) if not isinstance(d, torch.device) else d
        if d.type != "cuda":
            raise ValueError(f"B200Flat needs a CUDA device; got {d}")
        return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())

    def _store(self, emb_f32: torch.Tensor) -> torch.Tensor:
        """fp32 rows -> storage rows (optionally L2-normalising on device first)."""
        if self.normalize:
            return ops.normalize_rows(emb_f32, cast_dtype=None if self.dtype == torch.float32 else self.dtype)
        return emb_f32 if self.dtype == torch.float32 else emb_f32.to(self.dtype)

    def _bind(self, rows: torch.Tensor) -> None:
        """(Re)create the native handle over ``rows`` -- a view of ``_buf`` or a tensor of its own.  Cheap: the handle
        borrows the memory; only the TMA tensor maps are rebuilt."""
        if self._buf is None or rows.data_ptr() != self._buf.data_ptr():
            self._buf = rows
        self.shard = ops.FlatShard(rows, self.first_global_id if self._positions is None else 0)
        self.config["dimensions"] = int(rows.shape[1])

    @property
    def capacity(self) -> int:
        return 0 if self._buf is None else int(self._buf.shape[0])

    # ------------------------------------------------------------------ txtai ANN API
    def index(self, embeddings) -> None:
        """Build from float32 [N, dim] (already L2-normalised unless config normalize=True)."""
        emb = _as_cuda_f32(embeddings, self.device)
        rows = self._store(emb).contiguous()
        self._positions = None
        self._next_id = int(rows.shape[0])
        self._bind(rows)
        self.config["offset"] = self._next_id

    def append(self, embeddings) -> None:
        emb = _as_cuda_f32(embeddings, self.device)
        new = self._store(emb)
        if self.shard is None:
            return self.index(embeddings)
        if new.shape[1] != self.shard.dim:
            raise ValueError(f"dimension mismatch: index has {self.shard.dim}, got {new.shape[1]}")
        n, m = self.shard.n, int(new.shape[0])
        if n + m > self.capacity:
            # grow geometrically: ONE allocation + one device copy of the live rows, then appends are in place again
            cap = max(n + m, int(self.capacity * self.growth) + 1)
            buf = torch.empty((cap, self.shard.dim), dtype=self.dtype, device=self._buf.device)
            buf[:n].copy_(self._buf[:n])
            self._buf = buf
        self._buf[n:n + m].copy_(new)
        rows = self._buf[:n + m]
        if self._positions is not None:
            extra = torch.arange(self._next_id, self._next_id + new.shape[0], dtype=torch.int64, device=rows.device)
            self._positions = torch.cat([self._positions, extra])
        self._next_id += int(new.shape[0])
        self._bind(rows)
        self.config["offset"] = self._next_id

    def delete(self, ids: Sequence[int]) -> None:
        """Remove rows by ANN id; the ids of the remaining rows do not change."""
        if self.shard is None or len(ids) == 0:
            return
        dev = self.shard.device
        n = self.shard.n
        pos = self._positions if self._positions is not None else \
            torch.arange(self.first_global_id, self.first_global_id + n, dtype=torch.int64, device=dev)
        kill = torch.as_tensor(list(ids), dtype=torch.int64, device=dev)
        keep = ~torch.isin(pos, kill)
        self._positions = pos[keep]
        src = torch.nonzero(keep).squeeze(1)              # surviving rows, ascending: src[j] >= j
        n_keep = int(src.numel())
        first = int((src != torch.arange(n_keep, device=dev)).to(torch.int64).argmax().item()) if n_keep else 0
        if n_keep and bool((src[first:] != torch.arange(first, n_keep, device=dev)).any()):
            # compact IN PLACE, chunk by chunk: a chunk's sources lie at or behind its destination and behind every
            # earlier destination, so nothing is overwritten before it is read; the temporary is one chunk, not the index
            for c in range(first, n_keep, self.compact_chunk):
                e = min(c + self.compact_chunk, n_keep)
                self._buf[c:e] = self._buf.index_select(0, src[c:e])
        self._bind(self._buf[:n_keep])

    def count(self) -> int:
        return 0 if self.shard is None else self.shard.n

    def search_tensors(self, queries: torch.Tensor, limit: int, mode=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-resident search: float32 CUDA [B, dim] -> (scores [B,k], ids [B,k]) on device."""
        if self.shard is None:
            raise RuntimeError("index is empty: call index() or load() first")
        s, i = self.shard.search(queries, int(limit), self.mode if mode is None else mode)
        if self._positions is not None:
            valid = i >= 0
            i = torch.where(valid, self._positions[i.clamp_min(0)], i)
        return s, i

    def search(self, queries, limit: int, mode=None) -> List[List[Tuple[int, float]]]:
        """txtai ANN contract: ``[[(id, score), ...] per query]``, descending score,
        ties -> lower id; fewer than ``limit`` entries when the index is smaller."""
        if self.shard is None:
            raise RuntimeError("index is empty: call index() or load() first")
        use_host_call = (not isinstance(queries, torch.Tensor) or not queries.is_cuda) and self._positions is None \
            and int(limit) <= ops.K_CALL_MAX      # (k > 128 is composed on the device side: ops.FlatShard._search_wide)
        if use_host_call:
            q = queries if isinstance(queries, torch.Tensor) else \
                torch.from_numpy(np.ascontiguousarray(np.asarray(queries, dtype=np.float32)))
            q = q.to(torch.float32)
            q = (q.unsqueeze(0) if q.dim() == 1 else q).contiguous()
            if q.shape[1] != self.shard.dim:
                raise ValueError(f"queries must be [B, {self.shard.dim}]; got {tuple(q.shape)}")
            hs, hi = self.shard.search_host(q, int(limit), self.mode if mode is None else mode)
            s_np, i_np = hs.numpy(), hi.numpy()
        else:
            s, i = self.search_tensors(_as_cuda_f32(queries, self.device), limit, mode)
            s_np, i_np = s.cpu().numpy(), i.cpu().numpy()
        out = []
        for b in range(s_np.shape[0]):
            ids_b, sc_b = i_np[b].tolist(), s_np[b].tolist()
            out.append([(i_, s_) for i_, s_ in zip(ids_b, sc_b) if i_ >= 0])
        return out

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        """Raw row-major matrix + JSON header."""
        if self.shard is None:
            raise RuntimeError("nothing to save")
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        rows = self.shard.rows.cpu()
        raw = rows.view(torch.uint8) if rows.dtype == torch.float32 else rows.view(torch.int16)
        np.save(path + ".rows.npy", raw.numpy(), allow_pickle=False)
        meta = {"version": _FORMAT_VERSION, "n": self.shard.n, "dim": self.shard.dim,
                "dtype": {torch.float32: "fp32", torch.bfloat16: "bf16", torch.float16: "fp16"}[self.dtype],
                "first_global_id": self.first_global_id, "next_id": self._next_id,
                "positions": None if self._positions is None else self._positions.cpu().tolist()}
        with open(path + ".json", "w", encoding="utf-8") as f:
            json.dump(meta, f)

    def load(self, path: str, row_range: Optional[Tuple[int, int]] = None) -> None:
        """Restore; ``row_range=(lo, hi)`` memory-maps and uploads only that row block
        (per-rank sharded load)."""
        with open(path + ".json", "r", encoding="utf-8") as f:
            meta = json.load(f)
        self.dtype = ops.DTYPES[meta["dtype"]]
        raw = np.load(path + ".rows.npy", mmap_mode="r", allow_pickle=False)
        lo, hi = (0, meta["n"]) if row_range is None else row_range
        if row_range is not None and meta.get("positions") is not None:
            raise NotImplementedError("sharded load of an index with deleted rows")
        with warnings.catch_warnings():      # the memory map is read-only and the tensor is only read (uploaded below)
            warnings.simplefilter("ignore", UserWarning)
            block = torch.from_numpy(np.ascontiguousarray(raw[lo:hi]))
        if self.dtype == torch.float32:
            rows = block.view(torch.float32).reshape(hi - lo, meta["dim"])
        else:
            rows = block.view(self.dtype).reshape(hi - lo, meta["dim"])
        self.first_global_id = int(meta.get("first_global_id", 0)) + lo
        self._positions = None if meta.get("positions") is None else \
            torch.tensor(meta["positions"], dtype=torch.int64, device=self.device)
        self._next_id = int(meta.get("next_id", meta["n"]))
        self._bind(rows.to(self.device).contiguous())


class B200Sharded(B200Flat):
    """The same ANN contract over a ROW-SHARDED index: one process per GPU (``torchrun``), rank r of G holds the
    contiguous row block ``shard_bounds(N, G, r)`` of the document matrix; a search is the local scan + fused top-k,
    one exchange of the ``[B, k]`` candidates over NVLink and the merge-top-k kernel (``sharded.ShardedFlat``), and
    returns the identical global answer on every rank.  ANN ids are global row positions, as in ``B200Flat``.

    This is what lets ``Embeddings(**cfg)`` (heavy_ranker.py:78-83) mount an index that does not fit -- or should not
    be scanned by -- one GPU: ``Embeddings(path=..., content=True, shards=True)`` under ``torchrun``.  Every rank calls
    ``index`` / ``load`` / ``search`` with the same arguments (SPMD); ``index`` keeps only this rank's block.
    ``append`` / ``delete`` / ``save`` are single-index operations: build and save with ``B200Flat``, then ``load``
    here (each rank memory-maps its own row block)."""

    def __init__(self, config: Optional[dict] = None):
        super().__init__(config)
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("a sharded index needs torch.distributed to be initialised (one process per GPU, torchrun)")
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.exchange = str(self.config.get("exchange", "auto"))
        self.n_total = 0
        self.sharded = None

    def _wrap(self, rows: torch.Tensor, n_total: int) -> None:
        from .sharded import ShardedFlat

        self.n_total = int(n_total)
        self.sharded = ShardedFlat(rows, self.n_total, mode=self.mode, exchange=self.exchange)
        self.shard = self.sharded.shard
        self._buf = rows
        self.first_global_id = self.shard.first_global_id
        self._next_id = self.n_total
        self.config["dimensions"] = int(rows.shape[1])
        self.config["offset"] = self.n_total

    def index(self, embeddings) -> None:
        from .sharded import shard_bounds

        n_total = int(embeddings.shape[0]) if hasattr(embeddings, "shape") else len(embeddings)
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        block = embeddings[lo:hi]
        emb = _as_cuda_f32(block, self.device) if hi > lo else \
            torch.empty((0, int(embeddings.shape[1])), dtype=torch.float32, device=self.device)
        self._positions = None
        self._wrap(self._store(emb).contiguous(), n_total)

    def load(self, path: str, row_range: Optional[Tuple[int, int]] = None) -> None:
        from .sharded import shard_bounds

        if row_range is not None:
            raise ValueError("a sharded index chooses its own row block")
        with open(path + ".json", "r", encoding="utf-8") as f:
            n_total = int(json.load(f)["n"])
        super().load(path, row_range=shard_bounds(n_total, self.world, self.rank))
        self._wrap(self.shard.rows, n_total)

    def count(self) -> int:
        return self.n_total

    def search_tensors(self, queries: torch.Tensor, limit: int, mode=None) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.sharded is None:
            raise RuntimeError("index is empty: call index() or load() first")
        s, i = self.sharded.search(queries, int(limit), mode)
        return s.clone(), i.clone()      # the sharded index reuses its output buffers

    def search(self, queries, limit: int, mode=None) -> List[List[Tuple[int, float]]]:
        if self.sharded is None:
            raise RuntimeError("index is empty: call index() or load() first")
        s, i = self.search_tensors(_as_cuda_f32(queries, self.device), limit, mode)
        s_np, i_np = s.cpu().numpy(), i.cpu().numpy()
        return [[(i_, s_) for i_, s_ in zip(i_np[b].tolist(), s_np[b].tolist()) if i_ >= 0] for b in range(s_np.shape[0])]

    def append(self, embeddings) -> None:
        raise NotImplementedError("append to a row-sharded index: append to the single index and reload the shards")

    def delete(self, ids: Sequence[int]) -> None:
        raise NotImplementedError("delete from a row-sharded index: delete from the single index and reload the shards")

    def save(self, path: str) -> None:
        raise NotImplementedError("a row-sharded index is loaded from a saved B200Flat index, not saved itself")
