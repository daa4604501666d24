"""CPU oracle for the dense-retrieval hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
(``vietnamese_qa_system_b200``) never does.

PARITY UNPINNED: the reference has no tests/golden vectors and its arithmetic
lives in the un-vendored, uninstallable third-party ``txtai`` -> ``faiss-cpu``
(``/root/reference/requirements.txt:74``).  This module restates txtai's exact
(flat / NumPy-backend) semantics -- see ``oracle.c`` for the per-function
citations.  Two independent restatements are kept so they can check each other:

* ``liboracle.so`` (C, ``oracle.c``): canonical-order fp32 and fp64 "semantic"
  scoring, heap-free sorted-list top-k, merge, pooling, normalisation.
* the ``np_*`` functions below: the txtai NumPy backend almost literally
  (``np.dot`` + a stable descending sort), used to cross-check the C code and as
  the multi-threaded (BLAS) CPU baseline in ``bench.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.oracle_normalize_rows.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p]
        L.oracle_normalize_rows.restype = None
        L.oracle_mean_pool.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
        L.oracle_mean_pool.restype = None
        L.oracle_dot_canonical.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int]
        L.oracle_dot_canonical.restype = ctypes.c_float
        L.oracle_search.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int64, f32p, i64p]
        L.oracle_search.restype = None
        L.oracle_scores.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, f32p]
        L.oracle_scores.restype = None
        L.oracle_merge_topk.argtypes = [f32p, i64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        f32p, i64p]
        L.oracle_merge_topk.restype = None
        L.oracle_agree.argtypes = [ctypes.c_int64, ctypes.c_float, ctypes.c_int64, ctypes.c_float,
                                   ctypes.c_double]
        L.oracle_agree.restype = ctypes.c_int
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


CANONICAL, SEMANTIC = 0, 1


def elems_per_chunk(storage: str) -> int:
    """Elements per 16-byte chunk of the stored row (canonical order, oracle.c)."""
    return {"fp32": 4, "float32": 4, "bf16": 8, "bfloat16": 8, "fp16": 8, "float16": 8}[storage]


def normalize_rows(x) -> np.ndarray:
    x = _f32(x)
    x2 = x.reshape(-1, x.shape[-1])
    out = np.empty_like(x2)
    lib().oracle_normalize_rows(_fp(x2), x2.shape[0], x2.shape[1], _fp(out))
    return out.reshape(x.shape)


def mean_pool(hidden, mask, normalize: bool = True) -> np.ndarray:
    h = _f32(hidden)
    m = _f32(mask)
    b, s, d = h.shape
    out = np.empty((b, d), dtype=np.float32)
    lib().oracle_mean_pool(_fp(h), _fp(m), b, s, d, int(normalize), _fp(out))
    return out


def search(docs, queries, k: int, mode: int = CANONICAL, storage: str = "fp32", first_id: int = 0):
    """Exact flat top-k.  Returns (scores float32[b,k], ids int64[b,k])."""
    d_ = _f32(docs)
    q_ = _f32(queries).reshape(-1, d_.shape[1])
    b = q_.shape[0]
    sc = np.empty((b, k), dtype=np.float32)
    ids = np.empty((b, k), dtype=np.int64)
    lib().oracle_search(_fp(d_), d_.shape[0], d_.shape[1], _fp(q_), b, k, mode, elems_per_chunk(storage),
                        first_id, _fp(sc), _ip(ids))
    return sc, ids


def scores(docs, queries, mode: int = CANONICAL, storage: str = "fp32") -> np.ndarray:
    d_ = _f32(docs)
    q_ = _f32(queries).reshape(-1, d_.shape[1])
    out = np.empty((q_.shape[0], d_.shape[0]), dtype=np.float32)
    lib().oracle_scores(_fp(d_), d_.shape[0], d_.shape[1], _fp(q_), q_.shape[0], mode,
                        elems_per_chunk(storage), _fp(out))
    return out


def merge_topk(cand_scores, cand_ids, k_out: int):
    """cand_* : [lists, b, k_in]; ids < 0 are padding."""
    cs = _f32(cand_scores)
    ci = np.ascontiguousarray(np.asarray(cand_ids, dtype=np.int64))
    lists, b, k_in = cs.shape
    sc = np.empty((b, k_out), dtype=np.float32)
    ids = np.empty((b, k_out), dtype=np.int64)
    lib().oracle_merge_topk(_fp(cs), _ip(ci), lists, b, k_in, k_out, _fp(sc), _ip(ids))
    return sc, ids


def agree(uid_a: int, score_a: float, uid_b: int, score_b: float, threshold: float = 0.4) -> bool:
    return bool(lib().oracle_agree(int(uid_a), float(score_a), int(uid_b), float(score_b), float(threshold)))


# ----------------------------------------------------------------------------
# numpy restatement (txtai NumPy backend + pooling), independent of the C code
# ----------------------------------------------------------------------------

def np_normalize_rows(x) -> np.ndarray:
    """txtai: ``x /= np.linalg.norm(x, axis=1)[:, None]`` in fp32; zero rows stay zero."""
    x = np.array(x, dtype=np.float32, copy=True)
    nrm = np.linalg.norm(x, axis=-1, keepdims=True)
    np.divide(x, nrm, out=x, where=nrm > 0)
    x[np.broadcast_to(nrm == 0, x.shape)] = 0
    return x


def np_mean_pool(hidden, mask, normalize: bool = True) -> np.ndarray:
    """``sum(tokens*mask,1) / clamp(mask.sum(1), min=1e-9)`` then optional normalise."""
    h = np.asarray(hidden, dtype=np.float32)
    m = np.asarray(mask, dtype=np.float32)[..., None]
    pooled = (h * m).sum(axis=1) / np.maximum(m.sum(axis=1), 1e-9)
    pooled = pooled.astype(np.float32)
    return np_normalize_rows(pooled) if normalize else pooled


def np_search(docs, queries, k: int, first_id: int = 0):
    """``np.dot(queries, data.T)`` + stable descending sort (ties keep lower position)."""
    d_ = np.asarray(docs, dtype=np.float32)
    q_ = np.asarray(queries, dtype=np.float32).reshape(-1, d_.shape[1])
    s = q_ @ d_.T
    n = d_.shape[0]
    kk = min(k, n)
    out_s = np.full((q_.shape[0], k), -np.inf, dtype=np.float32)
    out_i = np.full((q_.shape[0], k), -1, dtype=np.int64)
    for b in range(q_.shape[0]):
        order = np.argsort(-s[b], kind="stable")[:kk]
        out_s[b, :kk] = s[b, order]
        out_i[b, :kk] = order + first_id
    return out_s, out_i


def np_search_fast(docs, queries, k: int, first_id: int = 0):
    """BLAS sgemm + argpartition (what faiss IndexFlatIP amounts to); bench CPU baseline.
    Tie order inside the partition is not defined -- not used for bit-exact checks."""
    d_ = np.asarray(docs, dtype=np.float32)
    q_ = np.asarray(queries, dtype=np.float32).reshape(-1, d_.shape[1])
    s = q_ @ d_.T
    kk = min(k, d_.shape[0])
    part = np.argpartition(-s, kk - 1, axis=1)[:, :kk]
    ps = np.take_along_axis(s, part, axis=1)
    order = np.lexsort((part, -ps), axis=1)
    ids = np.take_along_axis(part, order, axis=1) + first_id
    return np.take_along_axis(ps, order, axis=1), ids
