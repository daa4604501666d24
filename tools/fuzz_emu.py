"""TEST INFRASTRUCTURE: randomised differential run of the emulated CUDA-core kernels (tests/emu) against the oracle.
Random shapes for the fp32-verify search (all storage types), the BM25 leg, K1 and the tcgen05 kernels (smem-resident and
TMEM-resident queries, incl. the opt-in QS variants up to dim 1024 and both k > 32 reduce kernels); bit-exact / tolerance
checks as in tests/test_emu_kernels.py.  usage: python tools/fuzz_emu.py [seed] [seconds]   (CPU only, no GPU needed)"""
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
L.emu_search_tensor.argtypes = T._MMA_ARGS
L.emu_search_ts.argtypes = T._TS_ARGS
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 120
n_ok = 0
while time.time() < t_end:
    which = rng.integers(5)
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
    elif which in (3, 4):  # tcgen05 kernels on the host models of ptx.cuh
        kind = ["bf16", "f16"][rng.integers(2)]
        qs = int(which == 4 and rng.integers(2))     # opt-in QS variant of the TMEM-resident-query kernel (dims up to 1024)
        dim = int(rng.choice([64, 128, 192, 256, 384, 512, 768] + ([832, 1024] if qs else [])))
        os.environ["VQA_REDUCE_SELECT"] = str(int(qs or rng.integers(2)))   # k > 32: either reduce kernel
        n = int(rng.integers(1, 900)); sm = int(rng.choice([1, 2, 3, 7, 148]))
        kb = dim // 64
        docs = T._unit(rng, n, dim)
        if n > 4: docs[n // 2] = docs[0]
        raw, vals = T._to_storage(docs, kind)
        mc = int(rng.integers(2))          # thread-block clusters with TMA multicast
        if which == 3:
            ncol = int(rng.choice([16, 32, 64, 128])); b = int(rng.integers(1, 2 * ncol)); k = int(rng.choice([1, 3, 10, 32, 40]))
            kps = int(rng.choice([d for d in (1, 2, 3) if kb % d == 0])); stages = int(rng.integers(2, 7))
            if b > 4 * (ncol // 2): b = 4 * (ncol // 2)
            nq_cta = ncol // 2
            lists = 0 if (nq_cta <= 32 and k <= 32) else nq_cta * ((k + 31) // 32 * 32) * 8 + nq_cta * 8
            boxes = (227 * 1024 - (1024 + kb * ncol * 128 + 1024 + lists)) // 16384     # as plan_tensor sizes the ring
            stages = min(stages, boxes // kps)
            if stages < 2: continue
            q = T._unit(rng, b, dim)
            out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
            if os.environ.get("FUZZ_VERBOSE"): print("mma", kind, dim, n, b, k, sm, ncol, stages, kps, mc, flush=True)
            rc = L.emu_search_tensor(T.ptr(raw), int(kind == "bf16"), n, dim, T.ptr(q), b, k, 7, sm, ncol, stages, kps, mc, T.ptr(out_s), T.ptr(out_i))
            tag = ("mma", kind, dim, n, b, k, sm, ncol, stages, kps, mc); tol = 1e-5
        else:
            split = int(rng.integers(2)); k = int(rng.choice([1, 5, 10, 26] + ([40, 100] if qs else []))) if not split else int(rng.choice([3, 20, 32, 50]))
            b = int(rng.integers(1, 200)); kps = int(rng.choice([d for d in (1, 2, 3, 4) if kb % d == 0])); stages = int(rng.integers(2, 6))
            ks = int(rng.integers(max(0, kb - 12), kb + 1)) if qs else 0
            kscan = k if split else k + 6
            depth = 32 if kscan <= 32 else kscan + 32
            lrows = 64 if (kscan > 32 and (split or (qs and b <= 64))) else 128
            fixed = 1024 + ks * 16384 + 1024 + (2 * 64 * 64 * 4 if split else 0) + lrows * depth * 8    # ts_smem_bytes_rt without the ring
            stages = min(stages, ((227 * 1024 - fixed) // 8192) // kps)                    # as plan_ts sizes the ring
            if stages < 2: continue
            q = T._unit(rng, b, dim)
            out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
            if os.environ.get("FUZZ_VERBOSE"): print("ts", kind, dim, n, b, k, sm, split, stages, kps, mc, qs, ks, flush=True)
            rc = L.emu_search_ts(T.ptr(raw), int(kind == "bf16"), n, dim, T.ptr(q), b, k, 7, sm, split, 6, stages, kps, mc, qs, ks, T.ptr(out_s), T.ptr(out_i))
            tag = ("ts", kind, dim, n, b, k, sm, split, stages, kps, mc, qs, ks); tol = 1e-5 if split else 5e-7
        assert rc == 0, (L.emu_last_error(), tag)
        ws, wi = oracle.search(vals, q, k, oracle.SEMANTIC, {"bf16": "bf16", "f16": "fp16"}[kind], first_id=7)
        fin = wi >= 0
        assert np.array_equal(out_i >= 0, fin), tag
        rec = np.mean([len(set(out_i[r][fin[r]]) & set(wi[r][fin[r]])) / max(1, fin[r].sum()) for r in range(b)])
        assert rec >= 0.999 or (np.abs(np.sort(out_s[fin]) - np.sort(ws[fin])).max() <= 2e-6), (tag, rec)
        assert np.abs(out_s[fin] - ws[fin]).max() <= tol * max(1.0, np.abs(ws[fin]).max()) + 2e-6, tag
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
