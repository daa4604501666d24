"""TEST INFRASTRUCTURE: randomised differential run of the emulated CUDA-core kernels (tests/emu) against the oracle.
Random shapes for the fp32-verify search (all storage types), the BM25 leg and K1; bit-exact / tolerance checks as in
tests/test_emu_kernels.py.  usage: python tools/fuzz_emu.py [seed] [seconds]   (CPU only, no GPU needed)"""
import ctypes
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from oracle import sparse as osp
from tests.emu import build as emu_build
from tests import test_emu_kernels as T
from tests.golden import sparse_inputs as si

L = ctypes.CDLL(emu_build.build())
L.emu_last_error.restype = ctypes.c_char_p
c=ctypes; _vp,_i32,_i64,_dbl=c.c_void_p,c.c_int32,c.c_int64,c.c_double
L.emu_sparse_search.argtypes = [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _dbl, _i32, _vp, _vp]
L.emu_bm25_weights.argtypes = [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _dbl, _dbl, _dbl, _vp]
L.emu_pool_normalize.argtypes = [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp]
T._bind_search(L)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 120
n_ok = 0
while time.time() < t_end:
    which = rng.integers(3)
    if which == 0:   # verify search
        kind = ["f32", "bf16", "f16"][rng.integers(3)]
        dim = int(rng.choice([8, 16, 64, 128, 200, 384, 512, 768, 1024])) if kind == "f32" else int(rng.choice([8, 64, 128, 384, 768, 1024, 1032]))
        n = int(rng.integers(1, 1500)); b = int(rng.integers(1, 20)); k = int(rng.choice([1, 2, 5, 10, 32, 33, 100, 128])); sm = int(rng.choice([1, 2, 5, 148]))
        docs = T._unit(rng, n, dim); q = T._unit(rng, b, dim)
        if n > 4: docs[n // 2] = docs[0]
        raw, vals = T._to_storage(docs, kind); code = {"f32": 0, "bf16": 1, "f16": 2}[kind]
        out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
        rc = L.emu_search_stream(T.ptr(raw), code, n, dim, T.ptr(q), b, k, 7, sm, T.ptr(out_s), T.ptr(out_i))
        assert rc == 0, (L.emu_last_error(), kind, dim, n, b, k, sm)
        ws, wi = oracle.search(vals, q, k, oracle.CANONICAL, {"f32": "fp32", "bf16": "bf16", "f16": "fp16"}[kind], first_id=7)
        assert np.array_equal(out_i, wi) and np.array_equal(out_s.view(np.uint32), ws.view(np.uint32)), ("verify", kind, dim, n, b, k, sm)
    elif which == 1:  # sparse
        n_docs = int(rng.choice([1, 5, 300, 3000, 20000, 40000])); vocab = int(rng.choice([5, 50, 500]))
        docs = si.zipf_corpus(n_docs, vocab, seed=int(rng.integers(1 << 30)), min_len=0 if rng.random() < .3 else 2, max_len=int(rng.integers(3, 30)))
        if not any(docs): continue
        normalize = bool(rng.integers(2)); sm = int(rng.choice([1, 3, 148]))
        e = T.EmuBM25(L, docs, normalize, sm_count=sm); ref = osp.BM25(normalize=normalize).index(docs)
        nonempty = [d for d in docs if d]
        qs = si.queries_from(nonempty, int(rng.integers(1, 6)), seed=int(rng.integers(1 << 30)), vocab=vocab)
        limit = int(rng.choice([1, 3, 10, 40, 200]))
        got = e.search(qs, limit)
        for qq, g in zip(qs, got):
            assert g == ref.search(qq, limit), ("sparse", n_docs, vocab, normalize, sm, limit, qq)
    else:  # K1
        kind = ["f32", "bf16", "f16"][rng.integers(3)]
        dim = int(rng.choice([8, 64, 128, 384, 512, 768, 1024, 2048]))
        b = int(rng.integers(1, 5)); s = int(rng.integers(1, 140))
        raw, vals = T._to_storage(rng.standard_normal((b, s, dim)).astype(np.float32), kind)
        mask = (rng.random((b, s)) < rng.random()).astype(np.int64)
        out = np.empty((b, dim), np.float32)
        code = {"f32": 0, "bf16": 1, "f16": 2}[kind]
        rc = L.emu_pool_normalize(T.ptr(raw), code, T.ptr(mask), 3, b, s, dim, 1, T.ptr(out))
        assert rc == 0, (L.emu_last_error(), kind, dim, b, s)
        want = oracle.mean_pool(vals, mask, True)
        assert np.abs(out - want).max() <= 2e-5, ("k1", kind, dim, b, s, np.abs(out - want).max())
    n_ok += 1
print("fuzz ok:", n_ok, "cases")
