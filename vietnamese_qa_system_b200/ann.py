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
        self.shard = ops.FlatShard(rows, self.first_global_id if self._positions is None else 0)
        self.config["dimensions"] = int(rows.shape[1])

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
        rows = torch.cat([self.shard.rows, new], dim=0)
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
        self._bind(self.shard.rows[keep].contiguous())

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
        use_host_call = (not isinstance(queries, torch.Tensor) or not queries.is_cuda) and self._positions is None
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
