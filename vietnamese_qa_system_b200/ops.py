"""Tensor-level entry points over the C ABI.

PyTorch is used for device memory, streams and nothing else: every function here
hands raw device pointers (``tensor.data_ptr()``) and the current CUDA stream to
``libvqa_b200.so``.  No function computes anything in torch, and none falls back
to the CPU.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N

_TORCH_TO_VQA = {torch.float32: N.F32, torch.bfloat16: N.BF16, torch.float16: N.F16}
_MASK_TO_VQA = {torch.int64: N.I64, torch.int32: N.I32, torch.uint8: N.U8, torch.bool: N.U8, torch.float32: N.F32}
DTYPES = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
          "fp16": torch.float16, "float16": torch.float16}


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (this engine has no CPU path); got device {t.device}")


def mode_id(mode) -> int:
    if isinstance(mode, int):
        return mode
    try:
        return N.MODES[str(mode).lower()]
    except KeyError:
        raise ValueError(f"unknown search mode {mode!r}; expected one of {sorted(N.MODES)}") from None


K_CALL_MAX = 128   # largest k of one vqa_search call (consts.h kMaxK)
K_SEGMENT = 128    # candidates kept per row segment by the k > 128 composition


def wide_segments(n_rows: int, k: int, k_seg: int = K_SEGMENT, max_candidates: int = 8192) -> List[Tuple[int, int]]:
    """Row segments ``[(lo, hi), ...]`` (ascending, non-empty) for a top-k with ``k > 128`` (vqa_merge_segments,
    include/vqa.h): about ``k / 64`` of them, so that a segment is expected to hold half of the ``k_seg`` candidates
    it can report (on unclustered data the 128th is then ~8 standard deviations away and a second pass is rare; every
    segment search has a fixed cost, so fewer is faster); never more than ``max_candidates / k_seg``."""
    if n_rows <= 0:
        return []
    want = max(2, -(-int(k) // 64))
    nseg = max(1, min(want, max_candidates // k_seg, n_rows))
    cuts = [n_rows * i // nseg for i in range(nseg + 1)]
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


def split_saturated(bounds: Sequence[Tuple[int, int]], saturated: Sequence[int], k_seg: int = K_SEGMENT):
    """Second-pass segment list: every saturated segment with more than ``k_seg`` rows (it may hold more of the
    answer than it could report) is cut in half.  Returns ``(new_bounds, changed)``."""
    out, changed = [], False
    for (a, b), flag in zip(bounds, saturated):
        if flag and b - a > k_seg:
            m = (a + b) // 2
            out += [(a, m), (m, b)]
            changed = True
        else:
            out.append((a, b))
    return out, changed


class FlatShard:
    """One row shard of the document-embedding matrix bound to a native index handle.

    Replaces the faiss index object txtai keeps per ``Embeddings``
    (reference call sites: heavy_ranker.py:86,88 build; :91-94 load; :98,100 search).
    ``rows`` is borrowed by the native side; this object keeps it alive.
    """

    def __init__(self, rows: torch.Tensor, first_global_id: int = 0):
        _need_cuda(rows, "rows")
        if rows.dim() != 2:
            raise ValueError(f"rows must be [n, dim]; got shape {tuple(rows.shape)}")
        if rows.dtype not in _TORCH_TO_VQA:
            raise ValueError(f"rows dtype must be float32, bfloat16 or float16; got {rows.dtype}")
        if rows.stride(1) != 1 and rows.shape[0] > 0:
            raise ValueError("rows must be row-major (unit stride along dim)")
        self.rows = rows
        self.n, self.dim = int(rows.shape[0]), int(rows.shape[1])
        self.first_global_id = int(first_global_id)
        self.device = rows.device
        L = N.lib()
        h = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        N.check(L.vqa_index_create(ctypes.byref(h), self.n, self.dim, _TORCH_TO_VQA[rows.dtype], dev_index,
                                   self.first_global_id))
        self._h = h
        stride_bytes = (rows.stride(0) if self.n > 1 else self.dim) * rows.element_size()
        N.check(L.vqa_index_bind(self._h, ctypes.c_void_p(rows.data_ptr() if self.n else 0), self.n, stride_bytes))
        self._ws = {}
        self._segs: Dict[Tuple[int, int], "FlatShard"] = {}   # row segments of the k > 128 composition
        self._wide = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        lib = getattr(N, "_lib", None) if N is not None else None  # module globals may be gone at exit
        if h is not None and lib is not None:
            lib.vqa_index_destroy(h)
            self._h = None

    # -- kernel-selection knobs (vqa_tuning_t; benchmarks and tests only) --------
    def get_tuning(self) -> "N.Tuning":
        t = N.Tuning()
        N.check(N.lib().vqa_index_get_tuning(self._h, ctypes.byref(t)))
        return t

    def set_tuning(self, tuning: "Optional[N.Tuning]" = None, **knobs) -> "N.Tuning":
        """Store knobs in the native handle: ``shard.set_tuning(ts_qs=0, reduce_select=0)`` changes the named
        fields of the current tuning; ``set_tuning(N.tuning_default())`` replaces it.  The workspace cache is
        dropped (its size depends on ``ts_extra``).  Must not overlap searches on this shard."""
        t = tuning if tuning is not None else self.get_tuning()
        t.update(**knobs)
        N.check(N.lib().vqa_index_set_tuning(self._h, ctypes.byref(t)))
        self._ws.clear()
        self._segs.clear()
        return t

    # -- planning / workspace -------------------------------------------------
    def plan(self, n_queries: int, k: int, mode="fast") -> Tuple[int, int]:
        fam, nl = ctypes.c_int32(), ctypes.c_int32()
        N.check(N.lib().vqa_search_plan(self._h, n_queries, k, mode_id(mode), ctypes.byref(fam), ctypes.byref(nl)))
        return fam.value, nl.value

    def describe(self, n_queries: int, k: int, mode="fast") -> dict:
        """The planner's choice for this shard's shape and tuning (``vqa_plan_describe_tuned``): kernel family, queries
        per CTA, ring geometry, and for the tcgen05 families whether the scan screens (``split == 0``: storage-precision
        queries, ``kscan`` candidates kept, exact re-scoring in the reduce) or carries hi/lo query columns."""
        props = torch.cuda.get_device_properties(self.device)
        out = (ctypes.c_int32 * 16)()
        smem = ctypes.c_size_t()
        t = self.get_tuning()
        N.check(N.lib().vqa_plan_describe_tuned(self.n, self.dim, _TORCH_TO_VQA[self.rows.dtype], n_queries, k,
                                                mode_id(mode), props.multi_processor_count,
                                                int(getattr(props, "shared_memory_per_block_optin", 232448)),
                                                ctypes.byref(t), out, ctypes.byref(smem)))
        keys = ("family", "pass_nq", "passes", "groups", "stages", "kps", "ncol", "split", "qs", "ks", "kscan", "k_out",
                "rescore", "tmem_query_cols", "m64")
        d = dict(zip(keys, list(out)))
        d["smem"] = int(smem.value)
        return d

    def workspace(self, n_queries: int, k: int, mode: int) -> torch.Tensor:
        """Scratch (candidate lists, thresholds) of one search, cached per (B, k, CUDA stream): searches issued
        on different streams run concurrently and must not share candidate buffers (include/vqa.h, threading).
        Two host threads searching on the SAME stream must pass their own ``workspace=`` to ``search``."""
        key = (n_queries, k, _stream(self.device))
        ws = self._ws.get(key)
        if ws is None:
            need = ctypes.c_size_t()
            N.check(N.lib().vqa_workspace_bytes(self._h, n_queries, k, mode, ctypes.byref(need)))
            ws = torch.zeros(need.value, dtype=torch.uint8, device=self.device)   # zeroed once (include/vqa.h)
            self._ws[key] = ws
        return ws

    def workspace_bytes(self, n_queries: int, k: int, mode="fast") -> int:
        need = ctypes.c_size_t()
        N.check(N.lib().vqa_workspace_bytes(self._h, n_queries, k, mode_id(mode), ctypes.byref(need)))
        return int(need.value)

    # -- search ---------------------------------------------------------------
    def search(self, queries: torch.Tensor, k: int, mode="fast", out_scores: Optional[torch.Tensor] = None,
               out_ids: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
               reduce_stream: Optional[torch.cuda.Stream] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries: float32 [B, dim] CUDA, L2-normalised.  Returns (scores f32 [B,k], ids i64 [B,k]).
        ``workspace``: caller-owned uint8 scratch of ``workspace_bytes(B, k)`` (default: per-stream cache).
        ``reduce_stream``: ``vqa_search_2s`` -- the scan runs on the current stream, the candidate reduce on
        ``reduce_stream`` (the outputs are valid in THAT stream's order); pass a workspace of your own."""
        _need_cuda(queries, "queries")
        if queries.dtype != torch.float32:
            raise ValueError(f"queries must be float32; got {queries.dtype}")
        if queries.dim() == 1:
            queries = queries.unsqueeze(0)
        if queries.dim() != 2 or queries.shape[1] != self.dim:
            raise ValueError(f"queries must be [B, {self.dim}]; got {tuple(queries.shape)}")
        if queries.stride(1) != 1 or queries.stride(0) % 4 != 0 or queries.data_ptr() % 16 != 0:
            queries = queries.contiguous()
        b = int(queries.shape[0])
        m = mode_id(mode)
        if out_scores is None:
            out_scores = torch.empty((b, k), dtype=torch.float32, device=self.device)
        if out_ids is None:
            out_ids = torch.empty((b, k), dtype=torch.int64, device=self.device)
        if k > K_CALL_MAX:
            if workspace is not None or reduce_stream is not None:
                raise ValueError(f"k > {K_CALL_MAX} is composed from segment searches: no workspace= / reduce_stream=")
            return self._search_wide(queries, int(k), m, out_scores, out_ids)
        ws = workspace if workspace is not None else self.workspace(b, k, m)
        q_stride = queries.stride(0) if b > 1 else self.dim
        if reduce_stream is not None:
            N.check(N.lib().vqa_search_2s(self._h, ctypes.c_void_p(queries.data_ptr()), q_stride, b, k, m,
                                          ctypes.c_void_p(out_scores.data_ptr()), ctypes.c_void_p(out_ids.data_ptr()),
                                          ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                          ctypes.c_void_p(_stream(self.device)), ctypes.c_void_p(reduce_stream.cuda_stream)))
            return out_scores, out_ids
        N.check(N.lib().vqa_search(self._h, ctypes.c_void_p(queries.data_ptr()), q_stride, b, k, m,
                                   ctypes.c_void_p(out_scores.data_ptr()), ctypes.c_void_p(out_ids.data_ptr()),
                                   ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(_stream(self.device))))
        return out_scores, out_ids

    # -- k > 128: segment searches + vqa_merge_segments --------------------------
    def _segment(self, lo: int, hi: int) -> "FlatShard":
        seg = self._segs.get((lo, hi))
        if seg is None:
            seg = FlatShard(self.rows[lo:hi], self.first_global_id + lo)
            N.check(N.lib().vqa_index_set_tuning(seg._h, ctypes.byref(self.get_tuning())))
            self._segs[(lo, hi)] = seg
        return seg

    def _search_wide(self, queries: torch.Tensor, k: int, mode: int, out_scores: torch.Tensor, out_ids: torch.Tensor):
        """Top-k for 128 < k <= 1024 (txtai's hybrid search asks each leg for 10 x limit candidates,
        heavy_ranker.py:98,100 with limit > 12): the shard is cut into row segments (views of the same rows, one
        native handle each), every segment reports its own best 128 through ``vqa_search`` and
        ``vqa_merge_segments`` sorts the survivors.  The merge also reports which segments may hold more of the
        answer than 128; those are halved and searched again -- the flags are read on the host, so unlike
        ``k <= 128`` this call synchronises the stream at least once.  Same order as every other search: score
        descending, ties -> the lower id."""
        L = N.lib()
        kmax, cmax = ctypes.c_int32(), ctypes.c_int32()
        N.check(L.vqa_merge_segments_limits(ctypes.byref(kmax), ctypes.byref(cmax)))
        if k > kmax.value:
            raise ValueError(f"k must be in [1, {kmax.value}] (got {k})")
        b = int(queries.shape[0])
        dev, ks, cap = self.device, K_SEGMENT, cmax.value // K_SEGMENT
        bounds = wide_segments(self.n, k, ks, cmax.value)
        if not bounds:
            out_scores.fill_(float("-inf"))
            out_ids.fill_(-1)
            return out_scores, out_ids
        key = (b, _stream(dev))
        st = self._wide.get(key)
        if st is None:   # two candidate buffers (a second pass moves the kept segments' lists into their new slots)
            st = [(torch.empty((cap, b, ks), dtype=torch.float32, device=dev),
                   torch.empty((cap, b, ks), dtype=torch.int64, device=dev)) for _ in range(2)]
            st.append(torch.empty(cap, dtype=torch.int32, device=dev))
            self._wide = {key: st}
        sat = st[2]
        have: Dict[Tuple[int, int], int] = {}
        cur = 0
        while True:
            S, I = st[cur]
            pS, pI = st[cur ^ 1]
            for slot, (lo, hi) in enumerate(bounds):
                old = have.get((lo, hi))
                if old is not None:
                    S[slot].copy_(pS[old])
                    I[slot].copy_(pI[old])
                    continue
                seg, kk = self._segment(lo, hi), min(ks, hi - lo)
                if kk == ks:
                    seg.search(queries, ks, mode, out_scores=S[slot], out_ids=I[slot])
                else:            # fewer rows than list slots: the tail stays empty (-inf / -1)
                    S[slot].fill_(float("-inf"))
                    I[slot].fill_(-1)
                    ts, ti = seg.search(queries, kk, mode)
                    S[slot, :, :kk].copy_(ts)
                    I[slot, :, :kk].copy_(ti)
            merge_segments(S[:len(bounds)], I[:len(bounds)], k, out_scores, out_ids, sat)
            flags = sat[:len(bounds)].tolist()          # device -> host: synchronises the stream
            new_bounds, changed = split_saturated(bounds, flags, ks)
            if not changed:
                return out_scores, out_ids
            if len(new_bounds) > cap:
                raise RuntimeError(f"top-{k}: more than {cap} row segments would be needed (the answer is "
                                   f"concentrated in a few row ranges); search with k <= {K_CALL_MAX} instead")
            kept = set(new_bounds)
            have = {seg_b: i for i, seg_b in enumerate(bounds) if seg_b in kept}
            bounds, cur = new_bounds, cur ^ 1

    def search_host(self, queries_host: torch.Tensor, k: int, mode="fast"):
        """Host-buffer search through ``vqa_search_host``: pinned fp32 [B,dim] in,
        pinned (scores, ids) out; H2D + search + D2H + stream sync inside the call."""
        if queries_host.is_cuda or queries_host.dtype != torch.float32 or not queries_host.is_contiguous():
            raise ValueError("queries_host must be a contiguous float32 CPU tensor")
        b = int(queries_host.shape[0])
        m = mode_id(mode)
        key = ("host", b, k, _stream(self.device))
        st = self._ws.get(key)
        if st is None:
            need = ctypes.c_size_t()
            N.check(N.lib().vqa_search_host_staging_bytes(self._h, b, k, m, ctypes.byref(need)))
            st = (torch.zeros(need.value, dtype=torch.uint8, device=self.device),
                  torch.empty((b, k), dtype=torch.float32).pin_memory(),
                  torch.empty((b, k), dtype=torch.int64).pin_memory())
            self._ws[key] = st
        staging, hs, hi = st
        N.check(N.lib().vqa_search_host(self._h, ctypes.c_void_p(queries_host.data_ptr()), b, k, m,
                                        ctypes.c_void_p(hs.data_ptr()), ctypes.c_void_p(hi.data_ptr()),
                                        ctypes.c_void_p(staging.data_ptr()), staging.numel(),
                                        ctypes.c_void_p(_stream(self.device))))
        return hs, hi

    def search_host_async(self, queries_host: torch.Tensor, k: int, mode="fast", slot: int = 0):
        """``vqa_search_host_async``: enqueue H2D + search + D2H on the current stream and return at once.
        Returns ``(scores, ids, event)`` -- pinned host tensors that hold the result once ``event`` (recorded
        after the copies) has completed.  ``slot`` selects one of several independent staging / output buffer
        sets so that calls can be in flight together (a serving loop alternates slot 0 / 1).  ``queries_host``
        must be pinned and must not be modified until the event has completed."""
        if queries_host.is_cuda or queries_host.dtype != torch.float32 or not queries_host.is_contiguous():
            raise ValueError("queries_host must be a contiguous float32 CPU tensor")
        if not queries_host.is_pinned():
            raise ValueError("queries_host must be pinned (page-locked) for an asynchronous copy")
        b = int(queries_host.shape[0])
        m = mode_id(mode)
        key = ("host_async", b, k, int(slot))
        st = self._ws.get(key)
        if st is None:
            need = ctypes.c_size_t()
            N.check(N.lib().vqa_search_host_staging_bytes(self._h, b, k, m, ctypes.byref(need)))
            st = (torch.zeros(need.value, dtype=torch.uint8, device=self.device),
                  torch.empty((b, k), dtype=torch.float32).pin_memory(),
                  torch.empty((b, k), dtype=torch.int64).pin_memory())
            self._ws[key] = st
        staging, hs, hi = st
        N.check(N.lib().vqa_search_host_async(self._h, ctypes.c_void_p(queries_host.data_ptr()), b, k, m,
                                              ctypes.c_void_p(hs.data_ptr()), ctypes.c_void_p(hi.data_ptr()),
                                              ctypes.c_void_p(staging.data_ptr()), staging.numel(),
                                              ctypes.c_void_p(_stream(self.device))))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return hs, hi, ev


def merge_topk(cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int):
    """K4: cand_* [lists, B, k_in] (float32 / int64, CUDA) -> (scores [B,k], ids [B,k])."""
    _need_cuda(cand_scores, "cand_scores")
    _need_cuda(cand_ids, "cand_ids")
    if cand_scores.dtype != torch.float32 or cand_ids.dtype != torch.int64:
        raise ValueError("cand_scores must be float32 and cand_ids int64")
    if cand_scores.dim() != 3 or cand_scores.shape != cand_ids.shape:
        raise ValueError("cand_scores / cand_ids must both be [lists, B, k_in]")
    cand_scores, cand_ids = cand_scores.contiguous(), cand_ids.contiguous()
    lists, b, k_in = (int(x) for x in cand_scores.shape)
    dev = cand_scores.device
    out_s = torch.empty((b, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((b, k), dtype=torch.int64, device=dev)
    N.check(N.lib().vqa_merge_topk(ctypes.c_void_p(cand_scores.data_ptr()), ctypes.c_void_p(cand_ids.data_ptr()),
                                   lists, b, k_in, k, ctypes.c_void_p(out_s.data_ptr()),
                                   ctypes.c_void_p(out_i.data_ptr()), dev.index or 0, ctypes.c_void_p(_stream(dev))))
    return out_s, out_i


def merge_segments(seg_scores: torch.Tensor, seg_ids: torch.Tensor, k: int, out_scores: Optional[torch.Tensor] = None,
                   out_ids: Optional[torch.Tensor] = None, saturated: Optional[torch.Tensor] = None):
    """``vqa_merge_segments``: seg_* [segments, B, k_seg] (float32 / int64, CUDA; lists sorted, segments in ascending
    row order) -> ``(scores [B,k], ids [B,k], saturated int32 [segments])`` for ``k <= 1024``."""
    _need_cuda(seg_scores, "seg_scores")
    _need_cuda(seg_ids, "seg_ids")
    if seg_scores.dtype != torch.float32 or seg_ids.dtype != torch.int64:
        raise ValueError("seg_scores must be float32 and seg_ids int64")
    if seg_scores.dim() != 3 or seg_scores.shape != seg_ids.shape:
        raise ValueError("seg_scores / seg_ids must both be [segments, B, k_seg]")
    if not (seg_scores.is_contiguous() and seg_ids.is_contiguous()):
        raise ValueError("seg_scores / seg_ids must be contiguous")
    nseg, b, ks = (int(x) for x in seg_scores.shape)
    dev = seg_scores.device
    if out_scores is None:
        out_scores = torch.empty((b, k), dtype=torch.float32, device=dev)
    if out_ids is None:
        out_ids = torch.empty((b, k), dtype=torch.int64, device=dev)
    if saturated is None:
        saturated = torch.empty(nseg, dtype=torch.int32, device=dev)
    N.check(N.lib().vqa_merge_segments(ctypes.c_void_p(seg_scores.data_ptr()), ctypes.c_void_p(seg_ids.data_ptr()),
                                       nseg, b, ks, int(k), ctypes.c_void_p(out_scores.data_ptr()),
                                       ctypes.c_void_p(out_ids.data_ptr()), ctypes.c_void_p(saturated.data_ptr()),
                                       dev.index or 0, ctypes.c_void_p(_stream(dev))))
    return out_scores, out_ids, saturated


def merge_topk_packed(gathered: torch.Tensor, n_lists: int, n_queries: int, k: int, ids_offset: int,
                      out_scores: Optional[torch.Tensor] = None, out_ids: Optional[torch.Tensor] = None):
    """K4 over the all-gathered packed blocks: ``gathered`` is uint8 ``[n_lists * block]`` where each rank's
    block holds float32 ``[B,k]`` scores at byte 0 and int64 ``[B,k]`` ids at byte ``ids_offset``."""
    _need_cuda(gathered, "gathered")
    block = gathered.numel() // n_lists
    if gathered.dtype != torch.uint8 or block * n_lists != gathered.numel() or block % 8 or ids_offset % 8:
        raise ValueError("gathered must be uint8 [n_lists * block] with 8-byte aligned blocks")
    dev = gathered.device
    if out_scores is None:
        out_scores = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
    if out_ids is None:
        out_ids = torch.empty((n_queries, k), dtype=torch.int64, device=dev)
    base = gathered.data_ptr()
    N.check(N.lib().vqa_merge_topk_strided(ctypes.c_void_p(base), ctypes.c_void_p(base + ids_offset), block // 4,
                                           block // 8, n_lists, n_queries, k, k,
                                           ctypes.c_void_p(out_scores.data_ptr()), ctypes.c_void_p(out_ids.data_ptr()),
                                           dev.index or 0, ctypes.c_void_p(_stream(dev))))
    return out_scores, out_ids


class PeerExchange:
    """Symmetric (peer-mapped) gather buffers + flags for the NCCL-free exchange (``vqa_exchange_push`` /
    ``vqa_merge_topk_wait``).  Layout of every rank's symmetric allocation: two parities x ``world`` slots
    of ``block`` bytes, then two parities x ``world`` uint64 flags.  Creation is collective."""

    def __init__(self, block: int, rank: int, world: int, device: torch.device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.block, self.rank, self.world, self.device = block, rank, world, device
        self.data_bytes = 2 * world * block
        total = self.data_bytes + 2 * world * 8
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group=group if group is not None else dist.group.WORLD)
        self.peer_bases = [int(p) for p in self.handle.buffer_ptrs]
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # everyone's flags are zero before the first push
        self.epoch = 0
        self._slots = (ctypes.c_void_p * world)()
        self._flags = (ctypes.c_void_p * world)()

    def push_and_merge(self, local: torch.Tensor, b: int, k: int, ids_off: int, out_s: torch.Tensor,
                       out_i: torch.Tensor):
        self.epoch += 1
        par = self.epoch & 1
        w, blk = self.world, self.block
        for r in range(w):
            base = self.peer_bases[r]
            self._slots[r] = base + (par * w + self.rank) * blk
            self._flags[r] = base + self.data_bytes + (par * w + self.rank) * 8
        st = ctypes.c_void_p(_stream(self.device))
        dev = self.device.index or 0
        N.check(N.lib().vqa_exchange_push(ctypes.c_void_p(local.data_ptr()), blk, self._slots, self._flags, w,
                                          self.epoch, dev, st))
        mine = self.buf.data_ptr()
        gathered = mine + par * w * blk
        N.check(N.lib().vqa_merge_topk_wait(ctypes.c_void_p(gathered), ctypes.c_void_p(gathered + ids_off), blk // 4,
                                            blk // 8, w, b, k, k, ctypes.c_void_p(out_s.data_ptr()),
                                            ctypes.c_void_p(out_i.data_ptr()),
                                            ctypes.c_void_p(mine + self.data_bytes + par * w * 8), self.epoch, dev, st))
        return out_s, out_i


def pool_normalize(hidden: torch.Tensor, mask: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """K1: hidden [B,S,D] (f32/bf16/f16), mask [B,S] (int64/int32/bool/uint8/f32) -> float32 [B,D]."""
    _need_cuda(hidden, "hidden")
    _need_cuda(mask, "mask")
    if hidden.dim() != 3 or mask.dim() != 2 or mask.shape != hidden.shape[:2]:
        raise ValueError(f"hidden must be [B,S,D] and mask [B,S]; got {tuple(hidden.shape)} / {tuple(mask.shape)}")
    if hidden.dtype not in _TORCH_TO_VQA:
        raise ValueError(f"hidden dtype must be float32/bfloat16/float16; got {hidden.dtype}")
    if mask.dtype not in _MASK_TO_VQA:
        raise ValueError(f"mask dtype must be int64/int32/uint8/bool/float32; got {mask.dtype}")
    hidden, mask = hidden.contiguous(), mask.contiguous()
    b, s, d = (int(x) for x in hidden.shape)
    out = torch.empty((b, d), dtype=torch.float32, device=hidden.device)
    N.check(N.lib().vqa_pool_normalize(ctypes.c_void_p(hidden.data_ptr()), _TORCH_TO_VQA[hidden.dtype],
                                       ctypes.c_void_p(mask.data_ptr()), _MASK_TO_VQA[mask.dtype], b, s, d,
                                       int(normalize), ctypes.c_void_p(out.data_ptr()), hidden.device.index or 0,
                                       ctypes.c_void_p(_stream(hidden.device))))
    return out


def normalize_rows(x: torch.Tensor, cast_dtype: Optional[torch.dtype] = None, inplace: bool = False):
    """Row-wise fp32 L2 normalise on device.  Returns the fp32 result, or the cast copy
    (bf16/fp16 storage rows) when ``cast_dtype`` is given."""
    _need_cuda(x, "x")
    if x.dtype != torch.float32 or x.dim() != 2:
        raise ValueError("x must be a float32 [n, dim] tensor")
    x = x if x.is_contiguous() else x.contiguous()
    n, d = int(x.shape[0]), int(x.shape[1])
    out = x if inplace else torch.empty_like(x)
    cast = None
    cast_kind = 0
    if cast_dtype is not None and cast_dtype != torch.float32:
        if cast_dtype not in (torch.bfloat16, torch.float16):
            raise ValueError(f"cast_dtype must be bfloat16 or float16; got {cast_dtype}")
        cast = torch.empty((n, d), dtype=cast_dtype, device=x.device)
        cast_kind = _TORCH_TO_VQA[cast_dtype]
    N.check(N.lib().vqa_normalize_rows(ctypes.c_void_p(x.data_ptr()), d, n, d, ctypes.c_void_p(out.data_ptr()), d,
                                       ctypes.c_void_p(cast.data_ptr()) if cast is not None else None, cast_kind, d,
                                       x.device.index or 0, ctypes.c_void_p(_stream(x.device))))
    return cast if cast is not None else out


def agree(ids_a: torch.Tensor, scores_a: torch.Tensor, ids_b: torch.Tensor, scores_b: torch.Tensor,
          threshold: float = 0.4):
    """Batched two-index agreement rule (heavy_ranker.py:110).  Returns (accept bool[n], combined f32[n])."""
    for t, nme in ((ids_a, "ids_a"), (scores_a, "scores_a"), (ids_b, "ids_b"), (scores_b, "scores_b")):
        _need_cuda(t, nme)
    ids_a, ids_b = ids_a.contiguous().view(-1), ids_b.contiguous().view(-1)
    scores_a, scores_b = scores_a.contiguous().view(-1), scores_b.contiguous().view(-1)
    if ids_a.dtype != torch.int64 or ids_b.dtype != torch.int64 or scores_a.dtype != scores_b.dtype \
            or scores_a.dtype not in (torch.float32, torch.float64):
        raise ValueError("ids must be int64 and both score vectors float32 (dense) or float64 (hybrid)")
    n = ids_a.numel()
    if not (ids_b.numel() == scores_a.numel() == scores_b.numel() == n):
        raise ValueError("all inputs must have the same length")
    dev = ids_a.device
    acc = torch.empty(n, dtype=torch.uint8, device=dev)
    comb = torch.empty(n, dtype=scores_a.dtype, device=dev)
    fn = N.lib().vqa_agree if scores_a.dtype == torch.float32 else N.lib().vqa_agree_f64
    N.check(fn(ctypes.c_void_p(ids_a.data_ptr()), ctypes.c_void_p(scores_a.data_ptr()),
               ctypes.c_void_p(ids_b.data_ptr()), ctypes.c_void_p(scores_b.data_ptr()), n,
               float(threshold), ctypes.c_void_p(acc.data_ptr()), ctypes.c_void_p(comb.data_ptr()),
               dev.index or 0, ctypes.c_void_p(_stream(dev))))
    return acc.view(torch.bool), comb


def hybrid_fuse(dense_scores: torch.Tensor, dense_ids: torch.Tensor, sparse_scores: torch.Tensor,
                sparse_ids: torch.Tensor, limit: int, w_dense: float = 0.5, w_sparse: float = 0.5, rrf: bool = False):
    """Dense + sparse candidate fusion (txtai ``Search`` with ``hybrid=True``; heavy_ranker.py:78-83,98,100):
    dense [B,kd] float32 / int64, sparse [B,ks] float64 / int64 -> (scores float64 [B,limit], ids int64 [B,limit]).
    ``rrf=True``: reciprocal-rank fusion (txtai's rule for un-normalised sparse scores); a leg with weight <= 0
    is ignored in either fusion."""
    for t, nme in ((dense_scores, "dense_scores"), (dense_ids, "dense_ids"), (sparse_scores, "sparse_scores"),
                   (sparse_ids, "sparse_ids")):
        _need_cuda(t, nme)
    if dense_scores.dtype != torch.float32 or sparse_scores.dtype != torch.float64 \
            or dense_ids.dtype != torch.int64 or sparse_ids.dtype != torch.int64:
        raise ValueError("dense scores must be float32, sparse scores float64, ids int64")
    if dense_scores.dim() != 2 or dense_scores.shape != dense_ids.shape or sparse_scores.dim() != 2 \
            or sparse_scores.shape != sparse_ids.shape or dense_scores.shape[0] != sparse_scores.shape[0]:
        raise ValueError("expected dense [B,kd] and sparse [B,ks] score/id pairs")
    dense_scores, dense_ids = dense_scores.contiguous(), dense_ids.contiguous()
    sparse_scores, sparse_ids = sparse_scores.contiguous(), sparse_ids.contiguous()
    b, kd = (int(x) for x in dense_scores.shape)
    ks = int(sparse_scores.shape[1])
    dev = dense_scores.device
    out_s = torch.empty((b, limit), dtype=torch.float64, device=dev)
    out_i = torch.empty((b, limit), dtype=torch.int64, device=dev)
    if rrf:
        N.check(N.lib().vqa_hybrid_fuse_rrf(ctypes.c_void_p(dense_ids.data_ptr()), kd, ctypes.c_void_p(sparse_ids.data_ptr()),
                                            ks, b, float(w_dense), float(w_sparse), int(limit),
                                            ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
                                            dev.index or 0, ctypes.c_void_p(_stream(dev))))
        return out_s, out_i
    N.check(N.lib().vqa_hybrid_fuse(ctypes.c_void_p(dense_scores.data_ptr()), ctypes.c_void_p(dense_ids.data_ptr()), kd,
                                    ctypes.c_void_p(sparse_scores.data_ptr()), ctypes.c_void_p(sparse_ids.data_ptr()),
                                    ks, b, float(w_dense), float(w_sparse), int(limit),
                                    ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
                                    dev.index or 0, ctypes.c_void_p(_stream(dev))))
    return out_s, out_i
