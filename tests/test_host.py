"""Host-side logic that needs no GPU: shard arithmetic, the sqlite mirror of setup_db.py,
prompt-context rendering, config handling, and the N>1 exchange driven over gloo
(world_size 2) with the oracle standing in for the two device steps."""
import os
import sqlite3

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tests.conftest import unit_rows
from vietnamese_qa_system_b200 import db
from vietnamese_qa_system_b200.embeddings import Embeddings
from vietnamese_qa_system_b200.ranker import NO_DOCS_MESSAGE, straighten_docs
from vietnamese_qa_system_b200.sharded import ShardedSearch, exchange_candidates, shard_bounds


def test_shard_bounds_cover_and_are_contiguous():
    for n in (0, 1, 7, 1024, 10_000_000, 10_000_001):
        for g in (1, 2, 4, 8):
            spans = [shard_bounds(n, g, r) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            per = -(-n // g) if n else 0
            assert all(hi - lo <= per for lo, hi in spans)
    assert shard_bounds(10_000_000, 8, 7) == (8_750_000, 10_000_000)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


# ---- setup_db.py mirror: the reference's own demo (setup_db.py:135-161) as assertions ----
FAKE = [{"name": n, "email": e} for n, e in [
    ("John Smith", "john.smith@example.com"), ("Alice Johnson", "alice.johnson@example.com"),
    ("David Williams", "david.williams@example.com"), ("Emily Brown", "emily.brown@example.com"),
    ("Michael Davis", "michael.davis@example.com"), ("Sophia Wilson", "sophia.wilson@example.com"),
    ("Daniel Jones", "daniel.jones@example.com"), ("Olivia Miller", "olivia.miller@example.com"),
    ("William Taylor", "william.taylor@example.com"), ("Ava Anderson", "ava.anderson@example.com")]]


def test_setup_db_demo_roundtrip(tmp_path):
    path = db.setup_database("documents", table_names=["documents", "wiki", "usr_info"],
                             fields=["(id INTEGER PRIMARY KEY AUTOINCREMENT, doc TEXT, source TEXT)",
                                     "(id INTEGER PRIMARY KEY AUTOINCREMENT, wikidoc TEXT, header TEXT, source TEXT)",
                                     "(id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT, email TEXT)"],
                             database_dir=str(tmp_path), verbose=False)
    assert path == os.path.join(str(tmp_path), "documents.db")
    db.insert_data(path, table_name="usr_info", data=FAKE, verbose=False)
    rows = db.query(path, "SELECT * FROM usr_info", fetch_size="all")
    assert [r[0] for r in rows] == list(range(1, 11))            # AUTOINCREMENT ids 1..10 in insertion order
    assert rows[0] == (1, "John Smith", "john.smith@example.com")
    assert db.query(path, "SELECT * FROM usr_info", fetch_size=1) == rows[0]
    assert db.query(path, "SELECT * FROM usr_info", fetch_size=3) == rows[:3]
    with pytest.raises(ValueError):
        db.query(path, "SELECT * FROM usr_info", fetch_size=0)
    with pytest.raises(sqlite3.OperationalError):
        db.query(path, "SELECT * FROM nope")
    db.drop_tables(path, ["documents", "wiki", "usr_info"], verbose=False)
    with pytest.raises(sqlite3.OperationalError):
        db.query(path, "SELECT * FROM usr_info")


def test_connect_database_asserts_like_the_reference(tmp_path):
    with pytest.raises(AssertionError):
        db.connect_database(str(tmp_path / "missing.db"))
    bad = tmp_path / "file.txt"
    bad.write_text("x")
    with pytest.raises(AssertionError):
        db.connect_database(str(bad))
    with pytest.raises(AssertionError):
        db.setup_database("x", database_dir=str(tmp_path / "nodir"), verbose=False)


def test_fetch_docs_batched(tmp_path):
    path = db.setup_database("documents", database_dir=str(tmp_path), verbose=False)
    db.insert_data(path, "documents", [{"doc": f"passage {i}", "source": "s"} for i in range(2000)], verbose=False)
    got = db.fetch_docs(path, [1, 2000, 1500, 1, 99999])
    assert got == {1: "passage 0", 2000: "passage 1999", 1500: "passage 1499"}
    # same answer as the reference's one-connection-per-id loop (heavy_ranker.py:102-104)
    assert db.query(path, "SELECT doc FROM documents WHERE id = 1500", fetch_size=1)[0] == got[1500]
    assert db.fetch_docs(path, []) == {}


def test_straighten_docs_matches_reference_format():
    assert straighten_docs(["a", "b"]) == " [CTX0]: a [ECTX0]  [CTX1]: b [ECTX1] "
    assert straighten_docs([]) == f"[ERROR]{NO_DOCS_MESSAGE}[ERROR]"


def test_embeddings_config_surface_without_gpu():
    e = Embeddings(hybrid=True, content=True, path="sentence-transformers/paraphrase-multilingual-mpnet-base-v2")
    assert e.config["hybrid"] is True and e.content is True and e.count() == 0
    # hybrid=True configures the BM25 term index exactly as txtai does
    assert e.config["scoring"] == {"method": "bm25", "normalize": True, "terms": True}
    with pytest.raises(RuntimeError):              # empty index
        e.search("xin chào", 1)
    with pytest.raises(NotImplementedError):
        Embeddings(scoring="tfidf")
    e2 = Embeddings({"content": False}, dtype="fp32")
    assert e2.config["dtype"] == "fp32"
    with pytest.raises(RuntimeError):
        e2.search(np.zeros(8, np.float32), 1)      # empty index
    assert Embeddings._unpack({"id": 7, "text": "t"}, 0)[0] == 7
    assert Embeddings._unpack((3, "t", None), 0) == (3, "t", None)
    assert Embeddings._unpack("bare", 5)[0] == 5


# ---- N>1 path on CPU: gloo, world_size 2 ------------------------------------------------
def _worker(rank, world, port, n, d, b, k, seed, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
        docs[n - 1] = docs[0]                       # duplicate straddling the two shards
        q[0] = docs[0]
        lo, hi = shard_bounds(n, world, rank)

        def local(queries, kk):                     # oracle stands in for the device scan (tests only)
            s, i = oracle.search(docs[lo:hi], queries.numpy(), kk, first_id=lo)
            return torch.from_numpy(s), torch.from_numpy(i)

        def merge(gs, gi, kk):                      # oracle stands in for the K4 kernel (tests only)
            s, i = oracle.merge_topk(gs.numpy(), gi.numpy(), kk)
            return torch.from_numpy(s), torch.from_numpy(i)

        s, i = ShardedSearch(local, merge).search(torch.from_numpy(q), k)
        fs, fi = oracle.search(docs, q, k)
        ok = np.array_equal(i.numpy(), fi) and np.array_equal(s.numpy(), fs)
        gs, gi = exchange_candidates(torch.full((b, k), float(rank)), torch.full((b, k), rank, dtype=torch.int64))
        ok = ok and gs.shape == (world, b, k) and all(float(gs[r, 0, 0]) == r and int(gi[r, 0, 0]) == r
                                                        for r in range(world))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(1001, 10), (5, 10)])
def test_sharded_search_gloo_world2(n, k):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 64, 3, k, 11, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True


def test_heavy_ranker_build_and_load_orchestration(tmp_path):
    """heavy_ranker.py:70-94 as HeavyRanker.build / .load, with a recording stand-in for Embeddings (the
    device-backed class is exercised by the GPU tests)."""
    from vietnamese_qa_system_b200 import corpus
    from vietnamese_qa_system_b200.ranker import REFERENCE_INDEXES, HeavyRanker, load_passages

    path = str(tmp_path / "documents.db")
    docs = corpus.insert_doc(path, texts=["Hà_Nội là thủ_đô. " * 40, "Sông Hồng chảy qua Hà_Nội. " * 30])
    rows = load_passages(path)
    assert [r["id"] for r in rows] == list(range(1, len(docs) + 1)) and [r["text"] for r in rows] == docs
    assert set(rows[0]) == {"id", "text", "source"}                                   # heavy_ranker.py:76

    calls = []

    class Recorder:
        def __init__(self, **cfg):
            self.cfg = cfg
            calls.append(("init", cfg))

        def index(self, data):
            calls.append(("index", len(data)))

        def save(self, p):
            calls.append(("save", p))

        def load(self, p):
            calls.append(("load", p))

    hr = HeavyRanker.build(path, str(tmp_path / "ix"), embeddings_cls=Recorder, dtype="fp32",
                           transform={"mini_lm": "enc384", "mpnet": "enc768"})
    assert [c[0] for c in calls] == ["init", "index", "save", "init", "index", "save"]
    assert calls[0][1] == {**REFERENCE_INDEXES["mini_lm"], "dtype": "fp32", "transform": "enc384"}
    assert calls[3][1] == {**REFERENCE_INDEXES["mpnet"], "dtype": "fp32", "transform": "enc768"}
    assert calls[0][1]["hybrid"] is True and calls[0][1]["content"] is True          # :78-83
    assert calls[1] == ("index", len(docs)) and calls[2] == ("save", str(tmp_path / "ix" / "mini_lm"))
    assert hr.database_path == path and hr.threshold == 0.4
    calls.clear()
    HeavyRanker.load(path, str(tmp_path / "ix"), embeddings_cls=Recorder)
    assert calls == [("init", {}), ("load", str(tmp_path / "ix" / "mini_lm")),
                     ("init", {}), ("load", str(tmp_path / "ix" / "mpnet"))]


def test_wide_k_segment_plan():
    """Host logic of the k > 128 composition (ops.FlatShard._search_wide): segments tile the shard in ascending
    order, about k / 64 of them and never more than the merge kernel holds; only saturated segments that are
    larger than their candidate list are split."""
    from vietnamese_qa_system_b200.ops import K_SEGMENT, split_saturated, wide_segments
    for n, k in ((10_000_000, 1000), (1_250_000, 129), (131, 130), (50, 1024), (1, 500), (70_000, 1024)):
        b = wide_segments(n, k)
        assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(hi > lo for lo, hi in b) and len(b) * K_SEGMENT <= 8192
        assert len(b) == min(max(2, -(-k // 64)), n, 64)
    assert wide_segments(0, 300) == []
    b, changed = split_saturated([(0, 300), (300, 400), (400, 1000)], [1, 1, 0])
    assert changed and b == [(0, 150), (150, 300), (300, 400), (400, 1000)]     # 100 rows <= 128: nothing more there
    assert split_saturated([(0, 100), (100, 200)], [1, 1]) == ([(0, 100), (100, 200)], False)
