"""The product's CUDA-core kernels executed on the host by the fiber emulator (tests/emu/) and checked against
the oracle -- kernel LOGIC parity on a machine without a GPU (the GPU tests remain the parity tests proper).
The emulator is test infrastructure: built from generated copies of the kernel headers, never used by the
product.  Covered: sparse_scan_kernel (sorted-merge and dense-tile paths) + sparse_finish_kernel,
bm25_weights_kernel, hybrid_fuse_kernel, pool_normalize_warp_kernel (K1)."""
import ctypes

import numpy as np
import pytest

import oracle
from oracle import sparse as osp
from tests.emu import build as emu_build
from tests.golden import sparse_inputs as si
from vietnamese_qa_system_b200.scoring import BM25

c = ctypes
_vp, _i32, _i64, _dbl = c.c_void_p, c.c_int32, c.c_int64, c.c_double


_SCHEDULES = {"thread order": 0, "reverse order": 1, "random interleaving": 2}


@pytest.fixture(scope="module")
def _emu_lib():
    if not emu_build.toolchain_available():
        pytest.skip("g++ or the CUDA headers are not available: the kernel emulator cannot be built here")
    L = ctypes.CDLL(emu_build.build())
    L.emu_set_schedule.argtypes = [_i32, c.c_uint64]
    L.emu_last_error.restype = c.c_char_p
    L.emu_sparse_search.argtypes = [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _dbl, _i32,
                                    _vp, _vp]
    L.emu_bm25_weights.argtypes = [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _dbl, _dbl, _dbl, _vp]
    L.emu_hybrid_fuse.argtypes = [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _dbl, _dbl, _i32, _i32, _vp, _vp]
    L.emu_pool_normalize.argtypes = [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _vp]
    return L


@pytest.fixture(params=["thread order", "reverse order", "random interleaving"])
def emu(request, _emu_lib):
    """Every test runs under three fiber schedules: kernels whose result depended on which thread reaches a
    barrier-free region first (a missing __syncthreads, an unordered shared-memory update) would differ."""
    _emu_lib.emu_set_schedule(_SCHEDULES[request.param], 12345)
    return _emu_lib


@pytest.fixture(params=["thread order", "random interleaving"])
def emu2(request, _emu_lib):
    """The tensor-core kernel tests (the slowest on the emulator) run under two of the three schedules to keep the
    CPU suite to a few minutes; tools/fuzz_emu.py and the CUDA-core kernel tests cover all three."""
    _emu_lib.emu_set_schedule(_SCHEDULES[request.param], 12345)
    return _emu_lib


def ptr(a):
    return a.ctypes.data_as(_vp)


def ok(L, rc):
    assert rc == 0, L.emu_last_error().decode()


class EmuBM25:
    """The product's host half (tokenise, CSR, statistics, query planning) + the emulated device half."""

    def __init__(self, L, docs, normalize=True, sm_count=148):
        self.L, self.sm = L, sm_count
        self.bm = BM25({"method": "bm25", "terms": True, "normalize": normalize})
        self.bm.build_postings(docs)
        h = self.bm._host
        self.weights = np.zeros(max(len(h["docs"]), 1), np.float32)
        self.idf = np.ascontiguousarray(self.bm.idf_host)
        ok(L, L.emu_bm25_weights(ptr(h["offsets"]), len(self.idf), ptr(h["docs"]), ptr(h["freqs"]), len(h["docs"]),
                                 ptr(self.idf), ptr(h["lengths"]), self.bm.k1, self.bm.b, float(self.bm.avgdl),
                                 ptr(self.weights)))

    def search(self, queries, limit):
        bm, h = self.bm, self.bm._host
        q_terms, q_freqs, q_meta, kmax = bm.plan_queries(queries, limit)
        kmax = max(kmax, min(limit, bm.total))
        lim = min(limit, kmax)
        b = len(queries)
        out_s, out_i = np.empty((b, lim), np.float64), np.empty((b, lim), np.int64)
        normalize = bool(bm.normalize and bm.avgscore)
        ok(self.L, self.L.emu_sparse_search(ptr(h["offsets"]), ptr(h["docs"]), ptr(self.weights), bm.total,
                                            len(self.idf), ptr(q_terms), ptr(q_freqs), ptr(q_meta), q_terms.shape[1],
                                            b, kmax, lim, int(normalize), float(bm.avgscore or 0.0), self.sm,
                                            ptr(out_s), ptr(out_i)))
        return [[(int(p), float(s)) for p, s in zip(ir, sr) if p >= 0] for sr, ir in zip(out_s, out_i)]


def test_emulated_bm25_weights_bit_exact(emu):
    docs, _ = si.corpus_small()
    e = EmuBM25(emu, docs)
    ref = osp.BM25().index(docs)
    off = e.bm._host["offsets"]
    for term, tid in e.bm.vocab.items():
        _, want = ref.weights(term)
        assert np.array_equal(e.weights[off[tid]:off[tid + 1]].view(np.uint32), want.view(np.uint32)), term


@pytest.mark.parametrize("normalize", [False, True])
def test_emulated_sparse_search_small_corpus(emu, normalize):
    """One tile, one CTA per query: the sorted-merge path, zero-fill, common-term merge, normalisation."""
    docs, queries = si.corpus_small()
    e = EmuBM25(emu, docs, normalize)
    ref = osp.BM25(normalize=normalize).index(docs)
    queries = queries[:10] + [["unknown-token"], []]
    for limit in (1, 10, 100, 1000):
        got = e.search(queries, limit)
        for q, g in zip(queries, got):
            assert g == ref.search(q, limit), (q, limit)


def test_emulated_sparse_search_both_paths_many_ctas(emu):
    """40 k documents = 3 tiles; a small 'SM count' gives one CTA per query (3 tiles each: > 2048 postings of a
    common term -> dense tiles), a large one gives 3 CTAs per query (sorted merge where the range is sparse)."""
    n = 40_000
    docs = []
    for i in range(n):
        d = ["pad%d" % (i % 5)]
        if i % 2 == 0:
            d += ["all"] * (1 + i % 3)
        if i % 9 == 0:
            d.append("mid")
        if i in (77, 20_001, 39_999):
            d.append("needle")
        docs.append(d)
    ref = osp.BM25().index(docs)
    queries = [["needle", "all"], ["mid", "all"], ["all"], ["needle"], ["needle", "mid", "all", "all"], ["pad3", "mid"]]
    for sm in (1, 148):
        e = EmuBM25(emu, docs, sm_count=sm)
        for limit in (1, 10):
            got = e.search(queries, limit)
            for q, g in zip(queries, got):
                assert g == ref.search(q, limit), (sm, q, limit)


def test_emulated_hybrid_fuse(emu):
    rng = np.random.default_rng(5)
    b, kd, ks = 9, 10, 10
    ds = np.sort(rng.random((b, kd)).astype(np.float32), axis=1)[:, ::-1].copy()
    di = np.stack([rng.choice(40, kd, replace=False) for _ in range(b)]).astype(np.int64)
    ss = np.sort(rng.random((b, ks)), axis=1)[:, ::-1].copy()
    spi = np.stack([rng.choice(40, ks, replace=False) for _ in range(b)]).astype(np.int64)
    spi[2, 4:], ss[2, 4:] = -1, -np.inf
    ds[3, :], ss[3, :] = 0.5, 0.5
    # weighted score sum (normalised sparse scores) and reciprocal-rank fusion (raw BM25); weights 0 / 1 skip a leg
    for rrf in (0, 1):
        for limit, w in ((1, 0.5), (5, 0.7), (20, 0.0), (20, 1.0), (7, 0.25)):
            out_s, out_i = np.empty((b, limit), np.float64), np.empty((b, limit), np.int64)
            ok(emu, emu.emu_hybrid_fuse(ptr(ds), ptr(di), kd, ptr(ss), ptr(spi), ks, b, w, 1 - w, limit, rrf, ptr(out_s),
                                        ptr(out_i)))
            for r in range(b):
                dense = [(int(i), float(s)) for i, s in zip(di[r], ds[r]) if i >= 0]
                sparse = [(int(i), float(s)) for i, s in zip(spi[r], ss[r]) if i >= 0]
                got = [(int(i), float(s)) for i, s in zip(out_i[r], out_s[r]) if i >= 0]
                assert got == osp.hybrid(dense, sparse, limit, w, normalized=not rrf), (r, limit, w, rrf)


def _to_storage(x, kind):
    """fp32 -> (raw storage array, fp32 values it holds)."""
    if kind == "f32":
        return x.copy(), x
    if kind == "f16":
        h = x.astype(np.float16)
        return h, h.astype(np.float32)
    u = x.view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)            # round to nearest even -> bf16 bits
    return r, (r.astype(np.uint32) << 16).view(np.float32)


@pytest.mark.parametrize("kind,dim", [("bf16", 768), ("bf16", 384), ("bf16", 1024), ("f16", 768), ("f32", 768),
                                      ("f32", 384), ("f32", 512), ("f32", 1024), ("f16", 1024), ("bf16", 64)])
def test_emulated_k1_pool_normalize(emu, kind, dim):
    """K1 for every token-unroll the register budget picks (2 at 768 bf16, 3 at 384 bf16, 4 at 1024 bf16, ...),
    right- and middle-masked sequences, a long masked run inside the valid range, an all-padding row and a
    single-token row, integer / byte / fractional float masks; non-finite values under masked tokens must not leak
    (masked tokens are never loaded)."""
    run = emu.emu_pool_normalize
    rng = np.random.default_rng(dim)
    b, s = 6, 70
    x = rng.standard_normal((b, s, dim)).astype(np.float32)
    lens = [70, 1, 0, 33, 17, 64]
    mask = (np.arange(s)[None, :] < np.array(lens)[:, None]).astype(np.int64)
    mask[3, 5:9] = 0                                                          # holes inside the valid range
    mask[0, 16:48] = 0                                                        # a long masked run
    x[0, 20] = np.inf                                                         # ... under the mask: never read / never used
    x[3, 6] = np.nan
    raw, vals = _to_storage(x, kind)
    vals = np.where(mask[:, :, None] != 0, vals, 0)
    code = {"f32": 0, "bf16": 1, "f16": 2}[kind]
    for normalize in (1, 0):
        out = np.empty((b, dim), np.float32)
        ok(emu, run(ptr(raw), code, ptr(mask), 3, b, s, dim, normalize, ptr(out)))
        want = oracle.mean_pool(vals, mask, bool(normalize))
        assert np.abs(out - want).max() <= 1e-5 * max(1.0, np.abs(want).max())
        assert np.all(out[2] == 0)                                            # all-padding row
    m8 = mask.astype(np.uint8)
    out8 = np.empty((b, dim), np.float32)
    ok(emu, run(ptr(raw), code, ptr(m8), 5, b, s, dim, 1, ptr(out8)))
    assert np.array_equal(out8, out) or np.abs(out8 - oracle.mean_pool(vals, mask, True)).max() < 1e-5
    mf = mask.astype(np.float32) * 0.5                                        # fractional weights (F32 masks)
    outf = np.empty((b, dim), np.float32)
    ok(emu, run(ptr(raw), code, ptr(mf), 0, b, s, dim, 0, ptr(outf)))
    assert np.abs(outf - oracle.mean_pool(vals, mask, False)).max() <= 1e-5 * max(1.0, np.abs(want).max())


# ---- the fp32 verify kernel family (scan_topk_kernel + reduce) on the emulator ---------------------------
def _bind_search(L):
    L.emu_search_stream.argtypes = [_vp, _i32, _i64, _i32, _vp, _i32, _i32, _i64, _i32, _vp, _vp]
    L.emu_merge_topk.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]
    L.emu_normalize_rows.argtypes = [_vp, _i64, _i32, _vp, _vp, _i32]
    L.emu_agree.argtypes = [_vp, _vp, _vp, _vp, _i64, _dbl, _vp, _vp]
    return L


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("kind,dim,n,b,k,sm", [
    ("f32", 768, 700, 1, 10, 148), ("f32", 768, 900, 8, 5, 3), ("f32", 384, 1500, 3, 10, 2), ("f32", 200, 600, 2, 7, 4),
    ("bf16", 768, 1100, 5, 10, 2), ("f16", 1024, 520, 11, 3, 6), ("f32", 768, 37, 4, 64, 148), ("bf16", 64, 2100, 9, 40, 4),
])
def test_emulated_verify_search_is_bit_identical_to_the_canonical_oracle(emu, kind, dim, n, b, k, sm):
    """north_star: 'an fp32 verification mode gives bit-identical top-k ids (ties broken by lower doc_id)' --
    here ids AND score bits, for every storage type, unrolled and generic row lengths, 1..11 queries (1/2/4/8
    per pass, several passes), one or several CTAs, k up to and beyond the row count, planted duplicates."""
    L = _bind_search(emu)
    rng = np.random.default_rng(n + dim + b)
    docs = _unit(rng, n, dim)
    docs[n // 3] = docs[5]                                  # exact duplicates: equal scores, lower id first
    docs[n - 1] = docs[5]
    q = _unit(rng, b, dim)
    q[0] = docs[5]
    raw, vals = _to_storage(docs, kind)
    code = {"f32": 0, "bf16": 1, "f16": 2}[kind]
    out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
    ok(L, L.emu_search_stream(ptr(raw), code, n, dim, ptr(q), b, k, 1000, sm, ptr(out_s), ptr(out_i)))
    storage = {"f32": "fp32", "bf16": "bf16", "f16": "fp16"}[kind]
    want_s, want_i = oracle.search(vals, q, k, oracle.CANONICAL, storage, first_id=1000)
    assert np.array_equal(out_i, want_i)
    assert np.array_equal(out_s.view(np.uint32), want_s.view(np.uint32))
    top = out_i[0, :3].tolist()
    if k >= 3 and n > 10:
        assert top == sorted(top) and set(top) == {1005, 1000 + n // 3, 1000 + n - 1}


def test_emulated_merge_normalize_agree(emu):
    L = _bind_search(emu)
    rng = np.random.default_rng(3)
    lists, b, k = 5, 6, 10
    cs = -np.sort(-rng.random((lists, b, k)).astype(np.float32), axis=2)
    ci = rng.permutation(lists * b * k).reshape(lists, b, k).astype(np.int64)
    cs[1, :, 7:], ci[1, :, 7:] = -np.inf, -1                                 # padded tails
    cs[2, 0, :] = cs[3, 0, :]                                                # ties across lists -> lower id first
    out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
    ok(L, L.emu_merge_topk(ptr(cs), ptr(ci), lists, b, k, k, ptr(out_s), ptr(out_i)))
    ws, wi = oracle.merge_topk(cs, ci, k)
    assert np.array_equal(out_i, wi) and np.array_equal(out_s, ws)
    x = rng.standard_normal((9, 768)).astype(np.float32)
    x[4] = 0
    out = np.empty_like(x)
    cast = np.empty((9, 768), np.uint16)
    ok(L, L.emu_normalize_rows(ptr(x), 9, 768, ptr(out), ptr(cast), 1))
    want = oracle.normalize_rows(x)
    assert np.abs(out - want).max() < 1e-6 and np.all(out[4] == 0)
    assert np.array_equal(cast, _to_storage(out, "bf16")[0])                  # bf16 cast = round to nearest even
    ida = np.array([1, 2, 3, 4], np.int64)
    idb = np.array([1, 2, 9, 4], np.int64)
    sa = np.array([0.25, 0.2, 0.9, 0.125], np.float32)
    sb = np.array([0.2, 0.2, 0.9, 0.25], np.float32)
    acc, comb = np.empty(4, np.uint8), np.empty(4, np.float32)
    ok(L, L.emu_agree(ptr(ida), ptr(sa), ptr(idb), ptr(sb), 4, 0.4, ptr(acc), ptr(comb)))
    assert acc.tolist() == [int(oracle.agree(a, x_, b_, y)) for a, x_, b_, y in zip(ida, sa, idb, sb)] == [1, 1, 0, 0]


@pytest.mark.parametrize("select", ["0", "1"])
@pytest.mark.parametrize("kind", ["bf16", "f16"])
def test_emulated_reduce_with_exact_rescoring(emu, monkeypatch, kind, select):
    """The screen-then-rescore stage behind the large-batch tensor-core scan: the 16 best candidates by their
    (storage-precision) screen scores are re-scored exactly in fp32 against the stored rows, re-sorted and the
    best 10 emitted -- so a true neighbour that the screen ranked 11th..16th is recovered."""
    emu.emu_reduce_rescore.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _i32, _i32, _vp, _vp, _vp]
    monkeypatch.setenv("VQA_REDUCE_SELECT", select)   # "1": the CTA-per-query radix-select kernel re-scores (opt-in)
    rng = np.random.default_rng(8)
    n, dim, b, lists, k_in, k_out, k_final = 400, 768, 3, 5, 16, 16, 10
    raw, vals = _to_storage(_unit(rng, n, dim), kind)
    q = _unit(rng, b, dim)
    exact = vals.astype(np.float64) @ q.astype(np.float64).T                  # [n, b]
    screen = (exact + rng.normal(0, 2e-3, exact.shape)).astype(np.float32)    # perturbed order inside the top 16
    cand_s = np.full((lists, b, k_in), -np.inf, np.float32)
    cand_i = np.full((lists, b, k_in), 0xFFFFFFFF, np.uint32)
    per = n // lists
    for l in range(lists):                                                    # list l = rows [l*per, (l+1)*per)
        for j in range(b):
            ids = np.arange(l * per, (l + 1) * per)
            order = np.lexsort((ids, -screen[ids, j].astype(np.float64)))[:k_in]
            cand_s[l, j], cand_i[l, j] = screen[ids[order], j], ids[order]
    out_s, out_i = np.empty((b, k_final), np.float32), np.empty((b, k_final), np.int64)
    ok(emu, emu.emu_reduce_rescore(ptr(cand_s), ptr(cand_i), lists, b, k_in, k_out, k_final, 50, ptr(raw), dim,
                                   int(kind == "bf16"), ptr(q), ptr(out_s), ptr(out_i)))
    for j in range(b):
        allc = np.arange(n)
        top16 = allc[np.lexsort((allc, -screen[:, j].astype(np.float64)))[:k_out]]
        want = top16[np.lexsort((top16, -exact[top16, j]))[:k_final]]
        assert out_i[j].tolist() == (want + 50).tolist()
        assert np.abs(out_s[j] - exact[want, j]).max() < 5e-7


@pytest.mark.parametrize("lists,b,k_in,k_out,lmod", [
    (148, 2, 100, 100, 1),     # top-100 over the lists of 148 CTAs (BASELINE configs[3]'s k; hybrid search's 10 x limit)
    (7, 4, 128, 128, 1),       # the largest k
    (12, 6, 40, 33, 3),        # three query chunks side by side: query j lives in the lists l % 3 == j // 2
    (3, 2, 64, 64, 1),         # fewer valid candidates than k_out: padded with (-inf, -1)
])
def test_emulated_radix_select_reduce_equals_the_list_insertion_reduce(emu, monkeypatch, lists, b, k_in, k_out, lmod):
    """reduce_select_kernel (k > 32: 64-bit keys in shared memory, MSB radix select, bitonic sort of the survivors)
    returns exactly what reduce_topk_kernel returns -- ids, score bits, padding -- on unsorted candidate lists with
    heavy score ties (decided by the lower row), empty slots, +0.0 / -0.0 scores and negative scores; and both
    equal numpy's lexsort."""
    emu.emu_reduce_u32.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _vp]
    rng = np.random.default_rng(lists * 1000 + k_in)
    qpg = 2 if lmod > 1 else 1
    cand_s = np.full((lists, b, k_in), -np.inf, np.float32)
    cand_i = np.full((lists, b, k_in), 0xFFFFFFFF, np.uint32)
    for j in range(b):
        mine = [l for l in range(lists) if lmod == 1 or l % lmod == j // qpg]
        total = len(mine) * k_in
        n_valid = total if lists > 3 else k_out - 9                    # (3, 2, 64, 64): too few candidates
        rows = rng.permutation(1 << 20)[:total].astype(np.uint32)      # distinct rows, unsorted lists
        rows[:2] = [0, 0x7FFFFFFE]                                     # extreme ids
        sc = rng.choice(np.array([0.75, 0.5, 0.25, 0.0, -0.0, -0.125, 0.5000001], np.float32), total)
        sc = np.where(rng.random(total) < 0.5, sc, rng.normal(0.3, 0.2, total).astype(np.float32)).astype(np.float32)
        keep = rng.permutation(total) < n_valid
        flat_s = np.where(keep, sc, -np.inf).astype(np.float32)
        flat_i = np.where(keep, rows, 0xFFFFFFFF).astype(np.uint32)
        for t, l in enumerate(mine):
            cand_s[l, j], cand_i[l, j] = flat_s[t * k_in:(t + 1) * k_in], flat_i[t * k_in:(t + 1) * k_in]
    outs = []
    for select in ("0", "1"):
        monkeypatch.setenv("VQA_REDUCE_SELECT", select)
        out_s, out_i = np.empty((b, k_out), np.float32), np.empty((b, k_out), np.int64)
        tau_g = np.full(b, 77, np.uint64)
        ok(emu, emu.emu_reduce_u32(ptr(cand_s), ptr(cand_i), lists, b, k_in, k_out, 1000, lmod, qpg, ptr(tau_g),
                                   ptr(out_s), ptr(out_i)))
        assert not tau_g.any()                                          # shared-threshold slots cleared for the next search
        outs.append((out_s.copy(), out_i.copy()))
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32))   # score BITS, incl. the sign of zero
    for j in range(b):
        mine = [l for l in range(lists) if lmod == 1 or l % lmod == j // qpg]
        s = np.concatenate([cand_s[l, j] for l in mine])
        i = np.concatenate([cand_i[l, j] for l in mine]).astype(np.int64)
        s, i = s[i != 0xFFFFFFFF], i[i != 0xFFFFFFFF]
        order = np.lexsort((i, -s.astype(np.float64)))[:k_out]
        n = len(order)
        assert outs[1][1][j, :n].tolist() == (i[order] + 1000).tolist()
        assert np.array_equal(outs[1][0][j, :n], s[order])             # (-0.0 == 0.0 here; the bit check is above)
        assert np.all(outs[1][1][j, n:] == -1) and np.all(np.isneginf(outs[1][0][j, n:]))


@pytest.mark.parametrize("lists,b,k_in,k_out", [(148, 3, 10, 10), (148, 2, 16, 32), (33, 4, 32, 32), (5, 3, 7, 5)])
def test_emulated_warp_reduce_early_exit_is_exact(emu, monkeypatch, lists, b, k_in, k_out):
    """Opt-in VQA_REDUCE_EARLY=1: the k <= 32 reduce stops reading once a window holding one entry of every
    (best-first sorted) list offered nothing above the running k-th best.  Same ids and score bits as the full pass,
    and as numpy -- with ties, short lists and winners buried deep in one list."""
    emu.emu_reduce_u32.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _vp]
    rng = np.random.default_rng(lists + k_in)
    cand_s = np.full((lists, b, k_in), -np.inf, np.float32)
    cand_i = np.full((lists, b, k_in), 0xFFFFFFFF, np.uint32)
    for j in range(b):
        rows = rng.permutation(1 << 20)[:lists * k_in].astype(np.uint32).reshape(lists, k_in)
        sc = rng.choice(np.linspace(0, 1, 40, dtype=np.float32), (lists, k_in))        # heavy ties
        if j == 1:
            sc[lists // 2] = 2.0                                                       # every winner in ONE list
        fill = rng.integers(0, k_in + 1, lists)                                        # short lists (padding last)
        fill[lists // 2] = k_in
        for l in range(lists):
            order = np.lexsort((rows[l], -sc[l].astype(np.float64)))
            n = fill[l]
            cand_s[l, j, :n], cand_i[l, j, :n] = sc[l][order][:n], rows[l][order][:n]
    outs = []
    for early in ("0", "1"):
        monkeypatch.setenv("VQA_REDUCE_EARLY", early)
        out_s, out_i = np.empty((b, k_out), np.float32), np.empty((b, k_out), np.int64)
        ok(emu, emu.emu_reduce_u32(ptr(cand_s), ptr(cand_i), lists, b, k_in, k_out, 0, 1, 1, None, ptr(out_s), ptr(out_i)))
        outs.append((out_s, out_i))
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32))
    for j in range(b):
        s_, i_ = cand_s[:, j].ravel(), cand_i[:, j].ravel().astype(np.int64)
        s_, i_ = s_[i_ != 0xFFFFFFFF], i_[i_ != 0xFFFFFFFF]
        order = np.lexsort((i_, -s_.astype(np.float64)))[:k_out]
        assert outs[1][1][j, :len(order)].tolist() == i_[order].tolist()
        assert np.all(outs[1][1][j, len(order):] == -1)


def test_emulated_peer_memory_exchange_and_flag_waiting_merge(emu):
    """ShardedFlat(exchange="p2p") on one host: every 'rank' pushes its packed [scores | ids] block into its slot
    of every peer's gather buffer and publishes the epoch; each rank's merge kernel acquires the flags and merges
    the gathered blocks in place -- identical result on every rank, equal to the oracle merge."""
    emu.emu_exchange_push.argtypes = [_vp, c.c_size_t, _vp, _vp, _i32, c.c_uint64]
    emu.emu_merge_topk_wait.argtypes = [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, c.c_uint64]
    rng = np.random.default_rng(4)
    world, b, k, epoch = 4, 5, 10, 7
    ids_off = (b * k * 4 + 15) // 16 * 16
    block = (ids_off + b * k * 8 + 15) // 16 * 16
    scores = -np.sort(-rng.random((world, b, k)).astype(np.float32), axis=2)
    ids = rng.permutation(world * b * k).reshape(world, b, k).astype(np.int64)
    locals_ = []
    for r in range(world):
        blk = np.zeros(block, np.uint8)
        blk[:b * k * 4] = scores[r].view(np.uint8).ravel()
        blk[ids_off:ids_off + b * k * 8] = ids[r].view(np.uint8).ravel()
        locals_.append(blk)
    gather = [np.zeros(world * block, np.uint8) for _ in range(world)]       # rank p's gather buffer
    flags = [np.zeros(world, np.uint64) for _ in range(world)]               # rank p's flags, one per source rank
    for r in range(world):                                                   # rank r pushes to every peer p
        slots = (_vp * world)(*[gather[p].ctypes.data + r * block for p in range(world)])
        flg = (_vp * world)(*[flags[p].ctypes.data + r * 8 for p in range(world)])
        ok(emu, emu.emu_exchange_push(ptr(locals_[r]), block, slots, flg, world, epoch))
    assert all(f.tolist() == [epoch] * world for f in flags)
    ws, wi = oracle.merge_topk(scores, ids, k)
    for p in range(world):
        out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
        base = gather[p].ctypes.data
        ok(emu, emu.emu_merge_topk_wait(_vp(base), _vp(base + ids_off), block // 4, block // 8, world, b, k,
                                        ptr(out_s), ptr(out_i), ptr(flags[p]), epoch))
        assert np.array_equal(out_i, wi) and np.array_equal(out_s, ws)


# ---- the tcgen05 / TMA kernels on host models of the PTX wrappers (tests/emu/ptx_emu.cuh) --------------------
_MMA_ARGS = [_vp, _i32, _i64, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp]
_TS_ARGS = [_vp, _i32, _i64, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]


def _recall(got, want):
    k = want.shape[1]
    return float(np.mean([len(set(got[r]) & set(want[r])) / k for r in range(want.shape[0])]))


@pytest.mark.parametrize("kind,dim,n,b,k,sm,ncol,stages,kps,mc", [
    ("bf16", 128, 300, 3, 5, 2, 16, 4, 2, 0),       # NCOL 16: hi and lo in one TMEM load; last tile partly outside
    ("bf16", 768, 1000, 8, 10, 3, 16, 4, 2, 0),     # the headline shape in miniature (dim 768, top-10), 3 CTAs
    ("f16", 768, 700, 20, 10, 4, 64, 3, 2, 0),      # fp16 rows (scaled residual), 32 queries per CTA
    ("bf16", 256, 1100, 40, 10, 8, 32, 4, 1, 0),    # 3 query chunks side by side through L2 (n_groups = 3)
    ("bf16", 192, 90, 5, 32, 148, 16, 4, 3, 0),     # fewer rows than one tile, k = 32 (full register lists), kps = 3
    ("f16", 128, 600, 6, 40, 2, 16, 5, 2, 0),       # k > 32: CTA-shared sorted lists behind the spin lock
    ("bf16", 64, 2000, 70, 3, 4, 128, 6, 1, 0),     # NCOL 128: 64 queries per CTA, 2 chunks, shared-memory lists
    ("bf16", 256, 900, 30, 10, 8, 32, 4, 2, 1),     # cluster of 2: each CTA multicasts half of every box
    ("f16", 128, 700, 40, 5, 16, 32, 3, 1, 1),      # cluster of 4 (3 chunks + an empty one), quarter-box slices
])
def test_emulated_tcgen05_search_matches_the_oracle(emu2, kind, dim, n, b, k, sm, ncol, stages, kps, mc):
    """mma_topk_kernel (TMA producer / MMA issuer / TMEM epilogue with register top-k lists and GPU-wide
    thresholds) + the reduce, on the host models of mbarrier, TMA (128-byte swizzle; multicast across a cluster),
    tcgen05.mma / commit / ld and named barriers.  Bar of the GPU tests: recall against fp32 arithmetic on the stored rows, rank-wise scores
    within 1e-5 relative; here the ids are in fact identical, duplicates lower id first."""
    emu2.emu_search_tensor.argtypes = _MMA_ARGS
    rng = np.random.default_rng(dim + n + b)
    docs, q = _unit(rng, n, dim), _unit(rng, b, dim)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    raw, vals = _to_storage(docs, kind)
    out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
    ok(emu2, emu2.emu_search_tensor(ptr(raw), int(kind == "bf16"), n, dim, ptr(q), b, k, 100, sm, ncol, stages, kps, mc,
                                  ptr(out_s), ptr(out_i)))
    want_s, want_i = oracle.search(vals, q, k, oracle.SEMANTIC, {"bf16": "bf16", "f16": "fp16"}[kind], first_id=100)
    assert _recall(out_i, want_i) >= 0.999
    assert np.abs(out_s - want_s).max() <= 1e-5 * np.abs(want_s).max() + 2e-6
    assert out_i[0, :2].tolist() == [103, 100 + n // 2] and np.all(np.diff(out_s, axis=1) <= 0)


@pytest.mark.parametrize("kind,dim,n,b,k,sm,split,stages,kps,mc", [
    ("bf16", 128, 300, 40, 10, 2, 0, 4, 2, 0),      # screen (k + 6 candidates, 16-entry register lists) + exact re-score
    ("bf16", 768, 500, 70, 10, 3, 0, 4, 4, 0),      # dim 768: the query block fills 384 of the 512 TMEM columns
    ("f16", 256, 500, 150, 5, 4, 0, 3, 2, 0),       # two query chunks of 128 side by side through L2, fp16 rows
    ("bf16", 192, 300, 20, 30, 2, 1, 4, 3, 0),      # k = 30: hi + lo query rows (64 queries per CTA), 32-entry lists
    ("bf16", 128, 200, 10, 100, 2, 1, 4, 2, 0),     # k = 100: binary heaps in shared memory, 8-warp reduce
    ("bf16", 256, 600, 200, 10, 6, 0, 4, 2, 1),     # cluster of 2 with TMA multicast (the default for B > 128)
])
def test_emulated_tmem_resident_query_search_matches_the_oracle(emu2, kind, dim, n, b, k, sm, split, stages, kps, mc):
    """ts_topk_kernel (query block written to tensor memory with tcgen05.st and used as the MMA's A operand,
    thread-per-query-row epilogue with register lists / append buffers / heaps) + the reduce with the exact
    re-scoring stage, on the host models.  Screen mode returns the exact fp32 scores (abs err ~1e-7)."""
    emu2.emu_search_ts.argtypes = _TS_ARGS
    rng = np.random.default_rng(dim + n + b + k)
    docs, q = _unit(rng, n, dim), _unit(rng, b, dim)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    raw, vals = _to_storage(docs, kind)
    out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
    ok(emu2, emu2.emu_search_ts(ptr(raw), int(kind == "bf16"), n, dim, ptr(q), b, k, 100, sm, split, 6, stages, kps, mc,
                              0, 0, ptr(out_s), ptr(out_i)))
    want_s, want_i = oracle.search(vals, q, k, oracle.SEMANTIC, {"bf16": "bf16", "f16": "fp16"}[kind], first_id=100)
    assert _recall(out_i, want_i) >= 0.999
    assert np.abs(out_s - want_s).max() <= (5e-7 if not split else 1e-5)
    assert out_i[0, :2].tolist() == [103, 100 + n // 2] and np.all(np.diff(out_s, axis=1) <= 0)
    if not split:
        assert np.array_equal(out_i, want_i)


@pytest.mark.parametrize("kind,dim,n,b,k,sm,split,stages,kps,mc,ks", [
    ("f16", 1024, 400, 64, 100, 3, 0, 5, 2, 0, 4),   # BASELINE configs[3] in miniature: dim 1024 fp16, B = 64, top-100:
                                                     # 12 query blocks in TMEM + 4 in shared memory, heaps for 64 rows,
                                                     # big-k screen -> radix-select reduce re-scores the 128 best
    ("bf16", 1024, 300, 40, 10, 2, 0, 4, 4, 0, 4),   # dim 1024 bf16 top-10: register lists, warp reduce re-scores 32
    ("bf16", 1024, 300, 30, 40, 2, 1, 3, 2, 0, 4),   # dim 1024 hi/lo rows, k = 40 heaps (bf16 keeps hi/lo for big k)
    ("bf16", 768, 500, 130, 10, 4, 0, 4, 4, 0, 4),   # dim 768 with ks = 4: 8 blocks in TMEM -> 4 accumulator stages; 2 chunks
    ("f16", 768, 400, 200, 10, 6, 0, 4, 2, 1, 6),    # cluster of 2 with TMA multicast, ks = 6 (5 accumulator stages)
    ("bf16", 128, 200, 20, 10, 2, 0, 4, 2, 0, 2),    # every query block in shared memory (ks = dim / 64): nothing in TMEM
    ("f16", 832, 200, 10, 5, 2, 0, 4, 1, 0, 1),      # dim 832 = 13 blocks: ks = 1
])
def test_emulated_query_block_split_between_tmem_and_smem(emu2, monkeypatch, kind, dim, n, b, k, sm, split, stages, kps,
                                                         mc, ks):
    """ts_topk_kernel<.., QS = true> (opt-in, VQA_TS_QS=1): the last ks 64-column blocks of the query block are staged
    in shared memory (128-byte swizzle) and multiplied with the smem-A form of tcgen05.mma into the same accumulator
    as the TMEM-A blocks -- what makes dim 1024 fit.  Same bars as the all-TMEM kernel."""
    emu2.emu_search_ts.argtypes = _TS_ARGS
    monkeypatch.setenv("VQA_REDUCE_SELECT", "1")
    rng = np.random.default_rng(dim + n + b + k)
    docs, q = _unit(rng, n, dim), _unit(rng, b, dim)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    raw, vals = _to_storage(docs, kind)
    out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
    ok(emu2, emu2.emu_search_ts(ptr(raw), int(kind == "bf16"), n, dim, ptr(q), b, k, 100, sm, split, 6, stages, kps, mc,
                              1, ks, ptr(out_s), ptr(out_i)))
    want_s, want_i = oracle.search(vals, q, k, oracle.SEMANTIC, {"bf16": "bf16", "f16": "fp16"}[kind], first_id=100)
    assert _recall(out_i, want_i) >= 0.999
    assert np.abs(out_s - want_s).max() <= (5e-7 if not split else 1e-5)
    assert out_i[0, :2].tolist() == [103, 100 + n // 2] and np.all(np.diff(out_s, axis=1) <= 0)
    if not split:   # exact re-scoring: the oracle's ids, up to swaps of neighbours whose fp64 scores agree to 2e-7
        assert _recall(out_i, want_i) == 1.0
        swapped = out_i != want_i
        assert swapped.mean() < 0.01
        pos = {(r, int(i)): s_ for r in range(b) for i, s_ in zip(want_i[r], want_s[r])}
        assert all(abs(pos[(r, int(out_i[r, c]))] - want_s[r, c]) < 2e-7 for r, c in zip(*np.nonzero(swapped)))


@pytest.mark.parametrize("kind,dim,n,b,k,sm,ncol,stages,kps,mc", [
    ("bf16", 256, 3000, 8, 10, 12, 16, 4, 2, 0),       # 8 queries per CTA: one batch of 8 interleaved merges
    ("bf16", 64, 900, 32, 10, 6, 64, 3, 1, 0),         # B = 32 (the headline shape): four batches per warm-up tile
    ("f16", 128, 2100, 40, 5, 12, 32, 4, 1, 1),        # cluster of 4, three query chunks, the last one short
    ("bf16", 64, 700, 3, 32, 6, 16, 4, 1, 0),          # k = 32: the whole register list is the answer
    ("f16", 128, 900, 5, 1, 8, 16, 4, 2, 0),           # k = 1
    ("bf16", 128, 5000, 16, 10, 2, 32, 4, 2, 0),       # two CTAs, ~20 tiles each: warm-up merges, then single inserts
])
@pytest.mark.parametrize("seed,dyn", [("1", "1"), ("0", "0"), ("1", "0")])
def test_emulated_register_list_epilogue_seed_and_dynamic_tiles(emu2, monkeypatch, kind, dim, n, b, k, sm, ncol, stages, kps,
                                                                mc, seed, dyn):
    """The register-list epilogue of mma_topk_kernel with and without (a) the warm-up seed -- the first tile only
    publishes per-query maxima into k slots whose minimum bounds the k-th best before any list exists (75 us of
    warm-up merges on the B200 otherwise) -- and (b) the dynamic tile schedule (tickets of one counter instead of the
    static round-robin share).  Neither may change the answer: the oracle's ids (three-way tie at the top: lower id
    first), scores within fp32 rounding; slots and counter are left zeroed for the next launch."""
    monkeypatch.setenv("VQA_SEED", seed)
    monkeypatch.setenv("VQA_DYN_TILES", dyn)
    emu2.emu_search_tensor.argtypes = _MMA_ARGS
    rng = np.random.default_rng(dim + n + b + k)
    docs, q = _unit(rng, n, dim), _unit(rng, b, dim)
    docs[n // 2] = docs[3]
    docs[n - 1] = docs[3]                 # three-way tie at the top of query 0
    q[0] = docs[3]
    raw, vals = _to_storage(docs, kind)
    emu2.emu_events_read_reset.restype = c.c_longlong
    emu2.emu_events_read_reset()
    out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
    ok(emu2, emu2.emu_search_tensor(ptr(raw), int(kind == "bf16"), n, dim, ptr(q), b, k, 100, sm, ncol, stages, kps, mc,
                                    ptr(out_s), ptr(out_i)))
    events = emu2.emu_events_read_reset()
    want_s, want_i = oracle.search(vals, q, k, oracle.SEMANTIC, {"bf16": "bf16", "f16": "fp16"}[kind], first_id=100)
    assert _recall(out_i, want_i) >= 0.999
    assert np.abs(out_s - want_s).max() <= 1e-5
    assert out_i[0, :min(k, 3)].tolist() == [103, 100 + n // 2, 100 + n - 1][:min(k, 3)]
    assert np.all(np.diff(out_s, axis=1) <= 0)
    assert events > 0                     # list updates are instrumented


def test_emulated_segment_merge_composes_a_wide_top_k(emu):
    """vqa_merge_segments + the host-side segment logic of ops.FlatShard (k > 128, hybrid search with limit > 12):
    per-segment top-128 lists (here from the oracle, as vqa_search writes them) -> the kernel's sorted top-k and
    saturation flags -> saturated segments halved and merged again, until the answer equals the oracle's top-k
    bit for bit.  The planted cluster makes the first pass inexact, so the flags are what makes the result right."""
    from vietnamese_qa_system_b200.ops import K_SEGMENT, split_saturated, wide_segments
    L = emu
    L.emu_merge_segments.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]
    rng = np.random.default_rng(77)
    n, d, b, k, ks = 2600, 32, 3, 300, K_SEGMENT
    docs, q = _unit(rng, n, d), _unit(rng, b, d)
    near = q[0] + 0.05 * rng.standard_normal((400, d)).astype(np.float32)
    docs[1000:1400] = near / np.linalg.norm(near, axis=1, keepdims=True)     # 400 near-copies of query 0 in a row
    docs[1003] = docs[1001]
    docs[2500] = docs[1001]                                                   # exact ties -> lower id first
    want_s, want_i = oracle.search(docs, q, k)
    bounds, passes, first = wide_segments(n, k), 0, None
    assert len(bounds) == 5 and bounds[0] == (0, 520) and bounds[-1][1] == n
    while True:
        seg_s = np.full((len(bounds), b, ks), -np.inf, np.float32)
        seg_i = np.full((len(bounds), b, ks), -1, np.int64)
        for s_, (lo, hi) in enumerate(bounds):
            kk = min(ks, hi - lo)
            seg_s[s_, :, :kk], seg_i[s_, :, :kk] = oracle.search(docs[lo:hi], q, kk, first_id=lo)
        out_s, out_i = np.empty((b, k), np.float32), np.empty((b, k), np.int64)
        sat = np.full(len(bounds), 7, np.int32)
        ok(L, L.emu_merge_segments(ptr(seg_s), ptr(seg_i), len(bounds), b, ks, k, ptr(out_s), ptr(out_i), ptr(sat)))
        passes += 1
        first = out_i.copy() if first is None else first
        bounds, changed = split_saturated(bounds, sat.tolist())
        if not changed:
            break
    assert passes >= 2 and not np.array_equal(first, want_i)
    assert np.array_equal(out_i, want_i) and np.array_equal(out_s, want_s)
    # fewer candidates than k: the tail is empty; a single short segment; -0.0 ties with +0.0
    seg_s = np.array([[[0.5, -0.0, -np.inf]], [[0.75, 0.0, -1.0]]], np.float32)
    seg_i = np.array([[[4, 9, -1]], [[20, 21, 22]]], np.int64)
    out_s, out_i, sat = np.empty((1, 7), np.float32), np.empty((1, 7), np.int64), np.zeros(2, np.int32)
    ok(L, L.emu_merge_segments(ptr(seg_s), ptr(seg_i), 2, 1, 3, 7, ptr(out_s), ptr(out_i), ptr(sat)))
    assert out_i.tolist() == [[20, 4, 9, 21, 22, -1, -1]] and np.signbit(out_s[0, 2]) and not np.signbit(out_s[0, 3])
    assert np.all(np.isneginf(out_s[0, 5:])) and sat.tolist() == [0, 1]
