"""ctypes binding of libvqa_b200.so (include/vqa.h).

The library is the product's only compute path.  There is NO CPU fallback: if the
shared object is missing and cannot be built, or a compute entry point is called
without a CUDA device, this module raises -- it never routes anywhere else.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvqa_b200.so")

# vqa_status
OK, E_INVALID, E_CUDA, E_UNSUPPORTED, E_NOMEM = 0, -1, -2, -3, -4
# vqa_dtype
F32, BF16, F16, I64, I32, U8 = 0, 1, 2, 3, 4, 5
# vqa_mode
MODE_VERIFY, MODE_FAST, MODE_FAST_STREAM, MODE_FAST_TENSOR, MODE_FAST_TS, MODE_FAST_PAIR = 0, 1, 2, 3, 4, 5

MODES = {"verify": MODE_VERIFY, "fp32": MODE_VERIFY, "fast": MODE_FAST, "stream": MODE_FAST_STREAM,
         "tensor": MODE_FAST_TENSOR, "ts": MODE_FAST_TS, "pair": MODE_FAST_PAIR}

EXPORTS = [
    "vqa_version", "vqa_last_error", "vqa_device_count", "vqa_index_create", "vqa_index_bind",
    "vqa_index_destroy", "vqa_debug_timeline", "vqa_workspace_bytes", "vqa_search", "vqa_search_2s", "vqa_search_host_staging_bytes", "vqa_search_host_async",
    "vqa_search_host", "vqa_merge_topk", "vqa_merge_topk_strided", "vqa_exchange_push", "vqa_merge_topk_wait",
    "vqa_merge_segments", "vqa_merge_segments_limits",
    "vqa_pool_normalize", "vqa_normalize_rows", "vqa_agree",
    "vqa_search_plan", "vqa_plan_describe", "vqa_plan_describe_tuned",
    "vqa_tuning_default", "vqa_tuning_from_env", "vqa_index_set_tuning", "vqa_index_get_tuning",
    "vqa_sparse_limits", "vqa_sparse_create", "vqa_sparse_bind", "vqa_sparse_destroy", "vqa_bm25_weights",
    "vqa_sparse_workspace_bytes", "vqa_sparse_search", "vqa_hybrid_fuse", "vqa_hybrid_fuse_rrf", "vqa_agree_f64",
]

ABI_VERSION = 121  # VQA_VERSION this binding was written against (include/vqa.h)


class Tuning(ctypes.Structure):
    """vqa_tuning_t (include/vqa.h): the kernel-selection knobs stored in an index handle."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "size", "ts_extra", "ss_screen", "mma_kps", "mma_stages", "mma_groups", "mma_multicast", "ts_qs",
        "ts_ks", "ts_split", "ts_groups", "reduce_select", "reduce_early", "pdl_chain", "tma_l2promo", "tma_hint",
        "stream_max_b", "stream_min_mb", "pair", "dyn_tiles", "seed", "wide", "ts_m64", "smem_reserve_kb")]

    KNOBS = ("ts_extra", "ss_screen", "mma_kps", "mma_stages", "mma_groups", "mma_multicast", "ts_qs", "ts_ks",
             "ts_split", "ts_groups", "reduce_select", "reduce_early", "pdl_chain", "tma_l2promo", "tma_hint",
             "stream_max_b", "stream_min_mb", "pair", "dyn_tiles", "seed", "wide", "ts_m64", "smem_reserve_kb")

    def update(self, **knobs) -> "Tuning":
        for key, val in knobs.items():
            name = key.lower()
            if name.startswith("vqa_"):          # the environment spelling, VQA_TS_QS -> ts_qs
                name = name[4:]
            if name not in self.KNOBS:
                raise ValueError(f"unknown tuning knob {key!r}; expected one of {self.KNOBS}")
            setattr(self, name, int(val))
        return self

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n in self.KNOBS}


def tuning_default(from_env: bool = False) -> Tuning:
    t = Tuning()
    check((lib().vqa_tuning_from_env if from_env else lib().vqa_tuning_default)(ctypes.byref(t)))
    return t


_lib = None
_lock = threading.Lock()


class VqaError(RuntimeError):
    """CUDA / driver failure reported by the native library."""


def _bind(L: ctypes.CDLL) -> None:
    c = ctypes
    vp, i32, i64, sz = c.c_void_p, c.c_int32, c.c_int64, c.c_size_t
    L.vqa_version.restype = c.c_int
    L.vqa_version.argtypes = []
    L.vqa_last_error.restype = c.c_char_p
    L.vqa_last_error.argtypes = []
    L.vqa_device_count.restype = c.c_int
    L.vqa_device_count.argtypes = []
    L.vqa_index_create.restype = c.c_int
    L.vqa_index_create.argtypes = [c.POINTER(vp), i64, i32, i32, i32, i64]
    L.vqa_index_bind.restype = c.c_int
    L.vqa_index_bind.argtypes = [vp, vp, i64, i64]
    L.vqa_debug_timeline.restype = c.c_int
    L.vqa_debug_timeline.argtypes = [vp, vp, sz]
    L.vqa_index_destroy.restype = c.c_int
    L.vqa_index_destroy.argtypes = [vp]
    L.vqa_workspace_bytes.restype = c.c_int
    L.vqa_workspace_bytes.argtypes = [vp, i32, i32, i32, c.POINTER(sz)]
    L.vqa_search.restype = c.c_int
    L.vqa_search.argtypes = [vp, vp, i64, i32, i32, i32, vp, vp, vp, sz, vp]
    L.vqa_search_2s.restype = c.c_int
    L.vqa_search_2s.argtypes = [vp, vp, i64, i32, i32, i32, vp, vp, vp, sz, vp, vp]
    L.vqa_search_host_staging_bytes.restype = c.c_int
    L.vqa_search_host_staging_bytes.argtypes = [vp, i32, i32, i32, c.POINTER(sz)]
    L.vqa_search_host.restype = c.c_int
    L.vqa_search_host.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]
    L.vqa_search_host_async.restype = c.c_int
    L.vqa_search_host_async.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]
    L.vqa_merge_topk.restype = c.c_int
    L.vqa_merge_topk.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, i32, vp]
    L.vqa_merge_topk_strided.restype = c.c_int
    L.vqa_merge_topk_strided.argtypes = [vp, vp, i64, i64, i32, i32, i32, i32, vp, vp, i32, vp]
    L.vqa_exchange_push.restype = c.c_int
    L.vqa_exchange_push.argtypes = [vp, sz, c.POINTER(vp), c.POINTER(vp), i32, c.c_uint64, i32, vp]
    L.vqa_merge_topk_wait.restype = c.c_int
    L.vqa_merge_topk_wait.argtypes = [vp, vp, i64, i64, i32, i32, i32, i32, vp, vp, vp, c.c_uint64, i32, vp]
    L.vqa_merge_segments.restype = c.c_int
    L.vqa_merge_segments.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, vp]
    L.vqa_merge_segments_limits.restype = c.c_int
    L.vqa_merge_segments_limits.argtypes = [c.POINTER(i32), c.POINTER(i32)]
    L.vqa_pool_normalize.restype = c.c_int
    L.vqa_pool_normalize.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, vp, i32, vp]
    L.vqa_normalize_rows.restype = c.c_int
    L.vqa_normalize_rows.argtypes = [vp, i64, i64, i32, vp, i64, vp, i32, i64, i32, vp]
    L.vqa_agree.restype = c.c_int
    L.vqa_agree.argtypes = [vp, vp, vp, vp, i64, c.c_double, vp, vp, i32, vp]
    L.vqa_search_plan.restype = c.c_int
    L.vqa_search_plan.argtypes = [vp, i32, i32, i32, c.POINTER(i32), c.POINTER(i32)]
    L.vqa_plan_describe.restype = c.c_int
    L.vqa_plan_describe.argtypes = [i64, i32, i32, i32, i32, i32, i32, i32, c.POINTER(i32), c.POINTER(sz)]
    L.vqa_plan_describe_tuned.restype = c.c_int
    L.vqa_plan_describe_tuned.argtypes = [i64, i32, i32, i32, i32, i32, i32, i32, c.POINTER(Tuning), c.POINTER(i32),
                                          c.POINTER(sz)]
    L.vqa_tuning_default.restype = c.c_int
    L.vqa_tuning_default.argtypes = [c.POINTER(Tuning)]
    L.vqa_tuning_from_env.restype = c.c_int
    L.vqa_tuning_from_env.argtypes = [c.POINTER(Tuning)]
    L.vqa_index_set_tuning.restype = c.c_int
    L.vqa_index_set_tuning.argtypes = [vp, c.POINTER(Tuning)]
    L.vqa_index_get_tuning.restype = c.c_int
    L.vqa_index_get_tuning.argtypes = [vp, c.POINTER(Tuning)]
    L.vqa_sparse_limits.restype = c.c_int
    L.vqa_sparse_limits.argtypes = [c.POINTER(i32), c.POINTER(i32)]
    L.vqa_sparse_create.restype = c.c_int
    L.vqa_sparse_create.argtypes = [c.POINTER(vp), i64, i64, i64, i32]
    L.vqa_sparse_bind.restype = c.c_int
    L.vqa_sparse_bind.argtypes = [vp, vp, vp, vp]
    L.vqa_sparse_destroy.restype = c.c_int
    L.vqa_sparse_destroy.argtypes = [vp]
    L.vqa_bm25_weights.restype = c.c_int
    L.vqa_bm25_weights.argtypes = [vp, i64, vp, vp, i64, vp, vp, c.c_double, c.c_double, c.c_double, vp, i32, vp]
    L.vqa_sparse_workspace_bytes.restype = c.c_int
    L.vqa_sparse_workspace_bytes.argtypes = [vp, i32, i32, c.POINTER(sz)]
    L.vqa_sparse_search.restype = c.c_int
    L.vqa_sparse_search.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, c.c_double, vp, vp, vp, sz, vp]
    L.vqa_agree_f64.restype = c.c_int
    L.vqa_agree_f64.argtypes = [vp, vp, vp, vp, i64, c.c_double, vp, vp, i32, vp]
    L.vqa_hybrid_fuse.restype = c.c_int
    L.vqa_hybrid_fuse.argtypes = [vp, vp, i32, vp, vp, i32, i32, c.c_double, c.c_double, i32, vp, vp, i32, vp]
    L.vqa_hybrid_fuse_rrf.restype = c.c_int
    L.vqa_hybrid_fuse_rrf.argtypes = [vp, i32, vp, i32, i32, c.c_double, c.c_double, i32, vp, vp, i32, vp]


def lib() -> ctypes.CDLL:
    """Load the native library, (re)building it first when it is absent or older than its sources
    (content stamp, see build.build).  Fails loudly; a stale binary is never bound with new prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            from . import build as _build

            try:
                _build.build()  # no-op when the stamp matches csrc/ + include/vqa.h; needs nvcc otherwise
            except Exception as exc:  # noqa: BLE001
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing and cannot be built ({exc}). The CUDA extension is the only "
                        "compute path of this package (no CPU fallback).") from exc
                # an existing binary that could not be refreshed is accepted only if it speaks this ABI (checked below)
            try:
                L = ctypes.CDLL(LIB_PATH)
            except OSError as exc:  # pragma: no cover - depends on the box
                raise RuntimeError(
                    f"cannot load {LIB_PATH}: {exc}. The CUDA extension is the only compute path of this "
                    "package (no CPU fallback); build it with `python -m vietnamese_qa_system_b200.build`."
                ) from exc
            L.vqa_version.restype = ctypes.c_int
            if L.vqa_version() != ABI_VERSION:
                raise RuntimeError(f"{LIB_PATH} reports ABI version {L.vqa_version()}, this binding needs {ABI_VERSION}: "
                                   "rebuild with `python -m vietnamese_qa_system_b200.build --force`")
            _bind(L)
            _lib = L
    return _lib


def last_error() -> str:
    msg = lib().vqa_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int) -> None:
    """Map a vqa_status to the Python exception the reference-facing API documents."""
    if status == OK:
        return
    msg = last_error()
    if status == E_INVALID:
        raise ValueError(msg)
    if status == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status == E_NOMEM:
        raise MemoryError(msg)
    raise VqaError(msg)


def device_count() -> int:
    return int(lib().vqa_device_count())


def require_cuda() -> None:
    if device_count() == 0:
        raise VqaError("no CUDA device available: vietnamese_qa_system_b200 has no CPU fallback")
