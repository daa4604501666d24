"""The reference-facing surface: Embeddings (txtai call shapes used at heavy_ranker.py:78-101),
the ANN plugin B200Flat, and the batched heavy_ranker flow -- results against the oracle."""
import zlib

import numpy as np
import pytest
import torch

import oracle
from tests.conftest import unit_rows
from vietnamese_qa_system_b200 import B200Flat, Embeddings, HeavyRanker, db

pytestmark = pytest.mark.gpu


class FakeEncoder:
    """Deterministic text -> vector map standing in for the sentence-transformers model
    (weights are not available offline); hashes the text into a seeded Gaussian."""

    def __init__(self, dim):
        self.dim = dim

    def __call__(self, texts):
        out = np.empty((len(texts), self.dim), np.float32)
        for r, t in enumerate(texts):
            seed = zlib.crc32(t.encode("utf-8"))
            out[r] = np.random.default_rng(seed).standard_normal(self.dim)
        return out


def make_db(tmp_path, n):
    path = db.setup_database("documents", database_dir=str(tmp_path), verbose=False)
    db.insert_data(path, "documents", [{"doc": f"Đoạn văn số {i}", "source": f"wiki/{i}"} for i in range(n)],
                   verbose=False)
    return path


def test_heavy_ranker_call_shapes(tmp_path):
    """heavy_ranker.py:70-101 with the two indexes (D=384, D=768), content=True, top-1, dict results."""
    path = make_db(tmp_path, 300)
    data = db.query(path, "SELECT * FROM documents", fetch_size=50000)
    data_str = [{"id": row[0], "text": row[1], "source": row[2]} for row in data]       # :74-76
    encs = {384: FakeEncoder(384), 768: FakeEncoder(768)}
    idx = {}
    for dim, name in ((384, "mini_lm"), (768, "mpnet")):
        e = Embeddings(hybrid=False, content=True, transform=encs[dim], dtype="fp32")
        e.index(data_str)                                                               # :86,88
        e.save(str(tmp_path / "embeddings_index" / name))                               # :87,89
        e2 = Embeddings(transform=encs[dim])
        e2.load(str(tmp_path / "embeddings_index" / name))                              # :91-94
        idx[dim] = e2
    query_str = "Đoạn văn số 123"
    for dim in (384, 768):
        hit = idx[dim].search(query_str, 1)[0]                                          # :98,100
        assert set(hit) == {"id", "text", "score"}
        assert hit["id"] == 124 and hit["text"] == query_str                            # sqlite ids are 1-based
        assert abs(hit["score"] - 1.0) < 1e-6 and isinstance(hit["score"], float)
        text = db.query(path, f"SELECT doc FROM documents WHERE id = {hit['id']}", fetch_size=1)[0]
        assert text == query_str                                                        # :102-104
    # default limit is 3 (txtai) and results are descending
    res = idx[768].search(query_str)
    assert len(res) == 3 and res[0]["score"] >= res[1]["score"] >= res[2]["score"]
    # batched agreement rule :110
    ranked = HeavyRanker(idx[384], idx[768], database_path=path).rank([query_str, "không có trong kho"])
    assert ranked[0]["match"] is True and ranked[0]["id_a"] == 124 and abs(ranked[0]["score"] - 2.0) < 1e-5
    assert ranked[0]["doc_a"] == query_str
    a, b = ranked[1], ranked[1]
    assert a["match"] == oracle.agree(a["id_a"], a["score_a"], b["id_b"], b["score_b"])


def test_embeddings_vectors_tuples_match_oracle():
    rng = np.random.default_rng(0)
    docs, q = unit_rows(rng, 5000, 768), unit_rows(rng, 7, 768)
    e = Embeddings(content=False, dtype="fp32")
    e.index([(f"doc-{i}", docs[i], None) for i in range(len(docs))])     # (id, data, tags) form, string ids
    res = e.batchsearch(q, 5)
    stored = e.ann.shard.rows.cpu().numpy()
    qn = e.batchtransform(q).cpu().numpy()
    os_, oi = oracle.search(stored, qn, 5, oracle.CANONICAL, "fp32")
    for b in range(7):
        assert [r[0] for r in res[b]] == [f"doc-{j}" for j in oi[b]]
        assert [r[1] for r in res[b]] == [float(x) for x in os_[b]]
    assert e.count() == 5000 and len(e.search(q[0])) == 3


def test_embeddings_bf16_default_and_limit_larger_than_index():
    rng = np.random.default_rng(1)
    docs = unit_rows(rng, 4, 384)
    e = Embeddings()
    e.index(list(docs))
    assert e.ann.dtype == torch.bfloat16
    res = e.search(docs[2], 10)
    assert len(res) == 4 and res[0][0] == 2


def test_b200flat_ann_contract(tmp_path):
    rng = np.random.default_rng(2)
    docs, q = unit_rows(rng, 3000, 768), unit_rows(rng, 5, 768)
    ann = B200Flat({"dtype": "fp32"})
    ann.index(docs)
    out = ann.search(q, 4)
    os_, oi = oracle.search(docs, q, 4)
    assert [[i for i, _ in row] for row in out] == oi.tolist()
    assert [[s for _, s in row] for row in out] == [[float(x) for x in r] for r in os_]
    # limit > 128 (any limit up to 1024, as faiss answers it): composed from row-segment searches, host queries included
    wide = ann.search(q[:2], 300)
    ws_, wi = oracle.search(docs, q[:2], 300)
    assert [[i for i, _ in row] for row in wide] == wi.tolist()
    assert [[s for _, s in row] for row in wide] == [[float(x) for x in r] for r in ws_]
    # append keeps ids stable and appended rows are searchable
    extra = unit_rows(rng, 10, 768)
    ann.append(extra)
    assert ann.count() == 3010 and ann.search(extra[3:4], 1)[0][0][0] == 3003
    # delete removes rows without renumbering the rest
    top = ann.search(q[:1], 3)[0]
    ann.delete([top[0][0]])
    after = ann.search(q[:1], 2)[0]
    assert [i for i, _ in after] == [top[1][0], top[2][0]] and ann.count() == 3009
    # persistence round trip (+ sharded row-range load)
    ann2 = B200Flat({"dtype": "bf16"})
    ann2.index(docs)
    ann2.save(str(tmp_path / "ix"))
    ann3 = B200Flat()
    ann3.load(str(tmp_path / "ix"))
    assert ann3.search(q, 4) == ann2.search(q, 4)
    half = B200Flat()
    half.load(str(tmp_path / "ix"), row_range=(1500, 3000))
    assert half.count() == 1500 and half.first_global_id == 1500
    ids = [i for i, _ in half.search(q[:1], 5)[0]]
    assert all(1500 <= i < 3000 for i in ids)


def test_hf_encoder_to_search_end_to_end():
    """Config E shape in miniature: a real (randomly initialised) transformer encoder forward ->
    K1 pool+normalise -> search; checked against the oracle pooling of the same hidden states."""
    transformers = pytest.importorskip("transformers")
    from vietnamese_qa_system_b200.vectors import HFEncoder

    cfg = transformers.BertConfig(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                                  intermediate_size=256, max_position_embeddings=64)
    torch.manual_seed(0)
    model = transformers.BertModel(cfg)

    class Tok:
        def __call__(self, texts, padding=True, truncation=True, max_length=None, return_tensors="pt"):
            ids = [[1] + [3 + (ord(c) % 190) for c in t][:30] + [2] for t in texts]
            m = max(len(x) for x in ids)
            inp = torch.tensor([x + [0] * (m - len(x)) for x in ids])
            att = torch.tensor([[1] * len(x) + [0] * (m - len(x)) for x in ids])
            return {"input_ids": inp, "attention_mask": att}

    enc = HFEncoder("unused", model=model, tokenizer=Tok(), batch=4, dtype=torch.float32)
    texts = [f"câu hỏi {i} " + "x" * (i % 7) for i in range(10)]
    emb = enc(texts)
    assert emb.shape == (10, 128) and emb.is_cuda
    for j, t in enumerate(texts):
        hidden, mask = enc.hidden_states([t])
        ref = oracle.mean_pool(hidden.float().cpu().numpy(), mask.cpu().numpy(), True)
        assert np.abs(emb[j].cpu().numpy() - ref[0]).max() < 2e-5
    e = Embeddings(content=True, transform=enc, dtype="fp32")
    e.index([{"id": i + 1, "text": t} for i, t in enumerate(texts)])
    hit = e.search(texts[4], 1)[0]
    assert hit["id"] == 5 and hit["text"] == texts[4] and abs(hit["score"] - 1.0) < 1e-5


def test_append_is_in_place_and_delete_compacts_in_chunks():
    """SURVEY.md 8(f) rank 4 'without full rebuild': appends write behind the live rows of a capacity-reserved buffer
    (the buffer address only changes when the capacity is exhausted, geometrically), delete compacts in place chunk
    by chunk; ids stay stable and results equal a freshly built index over the surviving rows."""
    rng = np.random.default_rng(21)
    base = unit_rows(rng, 3000, 128)
    ann = B200Flat({"dtype": "fp32", "compact_chunk": 257})
    ann.index(base[:1000])
    ptrs = set()
    for lo in range(1000, 3000, 100):
        ann.append(base[lo:lo + 100])
        ptrs.add(ann._buf.data_ptr())
    assert ann.count() == 3000 and ann.capacity >= 3000
    assert len(ptrs) <= 4                                   # 20 appends, at most a few re-allocations (growth 1.5x)
    q = base[[5, 1500, 2999]] 
    assert [r[0][0] for r in ann.search(q, 1)] == [5, 1500, 2999]
    kill = [0, 5, 6, 7, 999, 1500, 2998] + list(range(2000, 2300))
    ann.delete(kill)
    assert ann.count() == 3000 - len(kill)
    keep = np.setdiff1d(np.arange(3000), kill)
    ref = B200Flat({"dtype": "fp32"})
    ref.index(base[keep])
    got = ann.search(q, 4)
    want = [[(int(keep[i]), s) for i, s in row] for row in ref.search(q, 4)]
    assert got == want                                      # same rows, original ids
    ann.append(base[:3])                                    # appended rows get fresh ids behind the old range
    assert ann.search(base[:1], 1)[0][0][0] == 3000
