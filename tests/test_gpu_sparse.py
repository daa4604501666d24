"""Sparse (BM25) leg, hybrid fusion and ``Embeddings(hybrid=True)`` on the GPU against the CPU oracle
(oracle/sparse.py).  Everything is bit-exact: positions identical, scores equal as Python floats."""
import ctypes
import zlib

import numpy as np
import pytest
import torch

import oracle
from oracle import sparse as osp
from tests.golden import sparse_inputs as si
from vietnamese_qa_system_b200 import Embeddings, HeavyRanker, _native as N, ops
from vietnamese_qa_system_b200.scoring import BM25

pytestmark = pytest.mark.gpu


def build(docs, normalize=True, **cfg):
    bm = BM25({"method": "bm25", "terms": True, "normalize": normalize, **cfg})
    bm.index(docs)
    ref = osp.BM25(normalize=normalize, k1=cfg.get("k1", 1.2), b=cfg.get("b", 0.75)).index(docs)
    return bm, ref


def assert_same(bm, ref, queries, limit):
    got = bm.batchsearch(queries, limit)
    for q, g in zip(queries, got):
        want = ref.search(q, limit)
        assert g == want, (q, limit, g[:3], want[:3])


def test_bm25_posting_weights_bit_exact():
    docs, _ = si.corpus_small()
    for cfg in ({}, {"k1": 0.9, "b": 0.4}):
        bm, ref = build(docs, **cfg)
        w = bm._dev["weights"].cpu().numpy()
        off = bm._host["offsets"]
        for term, tid in bm.vocab.items():
            uids, want = ref.weights(term)
            assert np.array_equal(w[off[tid]:off[tid + 1]].view(np.uint32), want.view(np.uint32)), term


@pytest.mark.parametrize("normalize", [False, True])
def test_sparse_search_small_corpus_all_limits(normalize):
    docs, queries = si.corpus_small()
    bm, ref = build(docs, normalize)
    for limit in (1, 3, 10, 30, 100, 400, 1000):       # 400 = N; 1000 > N
        assert_same(bm, ref, queries, limit)
    assert bm.search(["unknown-token"], 5) == []
    assert bm.search([], 5) == []


@pytest.mark.parametrize("batch", [1, 7, 64])
def test_sparse_search_many_tiles(batch):
    """70 k documents = 5 score tiles; the batch size changes how tiles are spread over CTAs."""
    docs = si.zipf_corpus(70_000, 5_000, seed=21, min_len=3, max_len=20)
    bm, ref = build(docs)
    queries = si.queries_from(docs, batch, seed=22 + batch, vocab=5_000)
    for limit in (1, 10, 50):
        assert_same(bm, ref, queries, limit)


def test_sparse_search_ties_return_lower_position_first():
    base = si.zipf_corpus(3_000, 800, seed=31)
    docs = base * 12                                     # every document 12 times: 36 k docs, 3 tiles
    bm, ref = build(docs, normalize=False)
    queries = [base[5][:3], base[77][:2], base[1234][:4] + ["w0"]]
    for limit in (5, 24, 40):
        assert_same(bm, ref, queries, limit)
    top = bm.search(queries[0], 12)
    assert len({s for _, s in top[:2]}) == 1 and top[0][0] < top[1][0]


def test_sparse_search_edge_documents():
    docs = [["alpha", "beta"], [], ["alpha"], ["gamma"] * 50, [], ["beta", "beta", "alpha"]]
    bm, ref = build(docs)
    for q in (["alpha"], ["beta", "alpha"], ["gamma", "gamma"], ["alpha", "nope"]):
        for limit in (1, 2, 6, 9):
            assert bm.search(q, limit) == ref.search(q, limit)
    with pytest.raises(ValueError):
        bm.search(["alpha"], 0)


def test_sparse_text_queries_use_the_tokenizer():
    texts = ["Hà Nội là thủ đô của Việt Nam", "Thành phố Hồ Chí Minh là thành phố lớn nhất",
             "Phở là món ăn nổi tiếng", "Sông Hồng chảy qua Hà Nội", "Vịnh Hạ Long ở Quảng Ninh"] * 3
    bm = BM25({"terms": True, "normalize": True})
    bm.index(texts)
    ref = osp.BM25().index([osp.tokenize(t) for t in texts])
    for q in ("thủ đô Hà Nội", "PHỞ!", "thành phố", "không có"):
        assert bm.search(q, 4) == ref.search(q, 4)


def test_sparse_save_load_round_trip(tmp_path):
    docs, queries = si.corpus_small()
    bm, ref = build(docs)
    bm.save(str(tmp_path / "scoring"))
    bm2 = BM25({"terms": True, "normalize": True})
    assert bm2.exists(str(tmp_path / "scoring"))
    bm2.load(str(tmp_path / "scoring"))
    assert bm2.avgscore == bm.avgscore and bm2.count() == 400
    assert bm2.batchsearch(queries, 7) == bm.batchsearch(queries, 7)


def test_hybrid_fuse_matches_python_fusion():
    rng = np.random.default_rng(5)
    b, kd, ks = 33, 10, 10
    ds = np.sort(rng.random((b, kd)).astype(np.float32), axis=1)[:, ::-1].copy()
    di = np.stack([rng.choice(60, kd, replace=False) for _ in range(b)]).astype(np.int64)
    ss = np.sort(rng.random((b, ks)), axis=1)[:, ::-1].copy()
    sparse_i = np.stack([rng.choice(60, ks, replace=False) for _ in range(b)]).astype(np.int64)
    for r in range(b):                                   # padding tails of different lengths, ties
        cut = int(rng.integers(0, ks + 1))
        sparse_i[r, cut:], ss[r, cut:] = -1, -np.inf
    di[3, 7:], ds[3, 7:] = -1, -np.inf
    ds[4, :] = 0.5
    ss[4, :] = np.where(sparse_i[4] >= 0, 0.5, -np.inf)
    for limit, w in ((1, 0.5), (3, 0.5), (10, 0.7), (20, 0.0), (5, 1.0)):
        fs, fi = ops.hybrid_fuse(torch.from_numpy(ds).cuda(), torch.from_numpy(di).cuda(),
                                 torch.from_numpy(ss).cuda(), torch.from_numpy(sparse_i).cuda(), limit, w, 1 - w)
        fs, fi = fs.cpu().numpy(), fi.cpu().numpy()
        for r in range(b):
            dense = [(int(i), float(s)) for i, s in zip(di[r], ds[r]) if i >= 0]
            sparse = [(int(i), float(s)) for i, s in zip(sparse_i[r], ss[r]) if i >= 0]
            want = osp.hybrid(dense, sparse, limit, w)
            got = [(int(i), float(s)) for i, s in zip(fi[r], fs[r]) if i >= 0]
            assert got == want, (r, limit, w)
            assert np.all(fi[r, len(want):] == -1) and np.all(np.isneginf(fs[r, len(want):]))


def test_agree_f64():
    ida = torch.tensor([1, 2, 3, 4], dtype=torch.int64, device="cuda")
    idb = torch.tensor([1, 2, 9, 4], dtype=torch.int64, device="cuda")
    sa = torch.tensor([0.2, 0.2, 0.9, 0.30000000000000004], dtype=torch.float64, device="cuda")
    sb = torch.tensor([0.2, 0.2000000000000001, 0.9, 0.1], dtype=torch.float64, device="cuda")
    acc, comb = ops.agree(ida, sa, idb, sb, 0.4)
    want = [a == b and x + y > 0.4 for a, b, x, y in zip(ida.tolist(), idb.tolist(), sa.tolist(), sb.tolist())]
    assert acc.cpu().tolist() == want == [False, True, False, False]
    assert comb.cpu().tolist() == [x + y for x, y in zip(sa.tolist(), sb.tolist())]


class FakeEncoder:
    def __init__(self, dim):
        self.dim = dim

    def __call__(self, texts):
        out = np.empty((len(texts), self.dim), np.float32)
        for r, t in enumerate(texts):
            out[r] = np.random.default_rng(zlib.crc32(t.encode("utf-8"))).standard_normal(self.dim)
        return out


def corpus_texts(n):
    words = ["hà", "nội", "sài", "gòn", "phở", "bún", "chả", "sông", "núi", "biển", "trường", "học", "sinh",
             "viên", "công", "nghệ", "máy", "tính", "ngôn", "ngữ", "lịch", "sử", "văn", "hóa", "kinh", "tế"]
    rng = np.random.default_rng(41)
    out = []
    for i in range(n):
        k = int(rng.integers(4, 14))
        out.append(" ".join(words[int(j)] for j in rng.integers(0, len(words), k)) + f" mã{i % 97}")
    return out


def oracle_hybrid(e, texts, queries, limit, weights=0.5, normalize=True):
    """Dense leg: canonical fp32 oracle on the index's own stored rows / normalised queries; sparse leg and
    fusion: oracle/sparse.py."""
    docs = e.ann.shard.rows.cpu().numpy()
    qs = e.batchtransform(queries).cpu().numpy()
    cand = limit * 10
    kd = min(cand, len(texts))
    s, i = oracle.search(docs, qs, kd, oracle.CANONICAL, "fp32")
    ref = osp.BM25(normalize=normalize).index([osp.tokenize(t) for t in texts])
    out = []
    for b, q in enumerate(queries):
        dense = [(int(p), float(x)) for p, x in zip(i[b], s[b]) if p >= 0]
        out.append(osp.hybrid(dense, ref.search(q, cand), limit, weights, normalized=normalize))
    return out


def test_embeddings_hybrid_matches_oracle(tmp_path):
    """The reference's configuration: Embeddings(hybrid=True, content=True, ...) -- heavy_ranker.py:78-83."""
    texts = corpus_texts(1500)
    enc = FakeEncoder(384)
    e = Embeddings(hybrid=True, content=True, transform=enc, dtype="fp32")      # fp32 -> verify mode: exact dense leg
    e.index([{"id": i + 1, "text": t, "source": f"s{i}"} for i, t in enumerate(texts)])
    queries = [texts[10], "phở bún chả", "mã5 hà nội", "không-có-từ-nào", texts[700] + " sông núi"]
    for limit, w in ((1, None), (3, None), (5, 0.7), (12, 0.2), (13, None), (20, 0.3)):   # > 12: dense k > 128
        got = e.batchsearch(queries, limit, w) if w is not None else e.batchsearch(queries, limit)
        want = oracle_hybrid(e, texts, queries, limit, 0.5 if w is None else w)
        for g, wv in zip(got, want):
            assert [(r["id"], r["score"]) for r in g] == [(p + 1, s) for p, s in wv]
            assert all(r["text"] == texts[r["id"] - 1] for r in g)
    hit = e.search(texts[10], 1)[0]                                          # the reference's call shape :98-101
    assert hit["id"] == 11 and set(hit) == {"id", "text", "score"}
    # persistence keeps the sparse leg (heavy_ranker.py:87-94)
    e.save(str(tmp_path / "mpnet"))
    e2 = Embeddings(transform=enc)
    e2.load(str(tmp_path / "mpnet"))
    assert e2.config["hybrid"] is True and e2.scoring is not None
    assert e2.batchsearch(queries, 3) == e.batchsearch(queries, 3)
    with pytest.raises(NotImplementedError):
        e.search(texts[0], 103)                                              # 1030 dense candidates > 1024
    with pytest.raises(ValueError):
        e.search(np.zeros(384, np.float32), 1)                               # the sparse leg needs text


def test_embeddings_hybrid_unnormalised_scoring_fuses_by_reciprocal_rank():
    """txtai fuses by weighted score sum only when the scoring index is normalised; `scoring="bm25"` next to a dense
    index (raw, unbounded BM25 scores) is fused by reciprocal rank, and weights 1 / 0 give the single-leg answer."""
    texts = corpus_texts(900)
    e = Embeddings(scoring="bm25", content=True, transform=FakeEncoder(384), dtype="fp32")
    e.index([{"id": i + 1, "text": t} for i, t in enumerate(texts)])
    assert e.scoring is not None and e.scoring.normalize is False and e.ann is not None
    queries = [texts[10], "phở bún chả", "mã5 hà nội", "không-có-từ-nào"]
    for limit, w in ((1, None), (4, 0.7), (10, 1.0), (10, 0.0), (3, [0.3, 0.3])):
        got = e.batchsearch(queries, limit, w) if w is not None else e.batchsearch(queries, limit)
        want = oracle_hybrid(e, texts, queries, limit, 0.5 if w is None else w, normalize=False)
        for g, wv in zip(got, want):
            assert [(r["id"], r["score"]) for r in g] == [(p + 1, s) for p, s in wv], (limit, w)
    h = Embeddings(hybrid=True, content=True, transform=FakeEncoder(384), dtype="fp32")   # normalised: weights 1 / 0
    h.index([{"id": i + 1, "text": t} for i, t in enumerate(texts)])
    for w in (1.0, 0.0):
        got = h.batchsearch(queries, 5, w)
        want = oracle_hybrid(h, texts, queries, 5, w)
        for g, wv in zip(got, want):
            assert [(r["id"], r["score"]) for r in g] == [(p + 1, s) for p, s in wv], w


def test_embeddings_keyword_only_index():
    texts = corpus_texts(300)
    e = Embeddings(keyword=True, content=True)
    e.index([{"id": 100 + i, "text": t} for i, t in enumerate(texts)])
    ref = osp.BM25(normalize=False).index([osp.tokenize(t) for t in texts])
    got = e.search("phở bún mã7", 4)
    want = ref.search("phở bún mã7", 4)
    assert [(r["id"], r["score"]) for r in got] == [(100 + p, s) for p, s in want]


def test_heavy_ranker_on_hybrid_indexes():
    texts = corpus_texts(400)
    a = Embeddings(hybrid=True, content=True, transform=FakeEncoder(384), dtype="fp32")
    b = Embeddings(hybrid=True, content=True, transform=FakeEncoder(768), dtype="fp32")
    docs = [{"id": i + 1, "text": t} for i, t in enumerate(texts)]
    a.index(docs)
    b.index(docs)
    queries = [texts[3], texts[250], "hoàn toàn không liên quan"]
    ranked = HeavyRanker(a, b).rank(queries)
    for q, r in zip(queries, ranked):
        ha, hb = a.search(q, 1)[0], b.search(q, 1)[0]
        assert (r["id_a"], r["score_a"], r["id_b"], r["score_b"]) == (ha["id"], ha["score"], hb["id"], hb["score"])
        assert r["match"] == (ha["id"] == hb["id"] and ha["score"] + hb["score"] > 0.4)
    assert ranked[0]["match"] and ranked[0]["id_a"] == 4


def test_sparse_c_abi_argument_errors():
    L = N.lib()
    h = ctypes.c_void_p()
    assert L.vqa_sparse_create(ctypes.byref(h), -1, 1, 1, 0) == N.E_INVALID
    assert L.vqa_sparse_create(ctypes.byref(h), 10, 1, 1, 99) == N.E_INVALID
    assert L.vqa_sparse_create(ctypes.byref(h), 10, 2, 3, 0) == N.OK
    need = ctypes.c_size_t()
    assert L.vqa_sparse_workspace_bytes(h, 0, 10, ctypes.byref(need)) == N.E_INVALID
    assert L.vqa_sparse_workspace_bytes(h, 4, 2000, ctypes.byref(need)) == N.E_INVALID
    assert L.vqa_sparse_workspace_bytes(h, 4, 10, ctypes.byref(need)) == N.OK and need.value == 4 * 10 * 8
    buf = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    p = ctypes.c_void_p(buf.data_ptr())
    # not bound yet
    assert L.vqa_sparse_search(h, p, p, p, 4, 4, 10, 5, 0, 0.0, p, p, p, 1024, None) == N.E_INVALID
    assert "vqa_sparse_bind" in N.last_error()
    assert L.vqa_sparse_bind(h, p, p, p) == N.OK
    assert L.vqa_sparse_search(h, p, p, p, 65, 4, 10, 5, 0, 0.0, p, p, p, 1024, None) == N.E_INVALID   # max_terms
    assert L.vqa_sparse_search(h, p, p, p, 4, 4, 10, 11, 0, 0.0, p, p, p, 1024, None) == N.E_INVALID   # limit > k_cand
    assert L.vqa_sparse_search(h, p, p, p, 4, 4, 10, 5, 1, 0.0, p, p, p, 1024, None) == N.E_INVALID    # normalize w/o avgscore
    assert L.vqa_sparse_search(h, p, p, p, 4, 4, 10, 5, 0, 0.0, p, p, p, 8, None) == N.E_NOMEM
    assert L.vqa_sparse_destroy(h) == N.OK
    assert L.vqa_hybrid_fuse(p, p, 0, p, p, 4, 1, 0.5, 0.5, 1, p, p, 0, None) == N.E_INVALID


def test_index_postings_equals_index():
    docs, queries = si.corpus_small()
    bm, _ = build(docs)
    h = bm._host
    bm2 = BM25({"method": "bm25", "terms": True, "normalize": True})
    bm2.index_postings(h["offsets"], h["docs"], h["freqs"], h["lengths"], list(bm.vocab))
    assert bm2.avgscore == bm.avgscore
    assert bm2.batchsearch(queries, 9) == bm.batchsearch(queries, 9)
    with pytest.raises(ValueError):
        bm2.index_postings(h["offsets"][:-1], h["docs"], h["freqs"], h["lengths"])


def test_sparse_both_accumulate_paths_and_zero_fill():
    """Crafted corpus that forces each stage-1 path: a CTA whose document range holds <= 2048 postings of
    the query merges them by sorting, otherwise it accumulates dense score tiles.  'needle' (3 documents) +
    the deferred common term 'all' needs the zero-fill candidates (positions < k_cand)."""
    n = 140_000
    docs = []
    for i in range(n):
        d = ["pad%d" % (i % 7)]
        if i % 12 == 0:
            d.append("mid")                   # 8.3 % of the documents: accumulated everywhere, ~2.7 k per 32 k range
        if i % 2 == 0:
            d += ["all"] * (1 + i % 3)        # 50 %: deferred common term, varying frequency
        if i in (77, 50_001, 139_999):
            d.append("needle")
        docs.append(d)
    bm, ref = build(docs)
    queries = [["needle", "all"], ["mid", "all"], ["needle", "mid", "all", "all"], ["mid"], ["all"], ["needle"],
               ["pad3", "mid"], ["nothing"]]
    for batch_mult in (1, 8):                  # 8 queries: 1 tile per CTA (sorted merge); 64: 2 tiles per CTA (tiles)
        qs = queries * batch_mult
        for limit in (1, 10, 30):
            assert_same(bm, ref, qs, limit)
    r = bm.search(["needle", "all"], 10)
    assert [p for p, _ in r[:3]] == sorted([77, 50_001, 139_999], key=lambda x: -dict(r)[x])
    assert len(r) == 10                        # 3 needle documents + zero-fill positions lifted by 'all'
