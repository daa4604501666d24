"""The reference's retriever script (inference_pipeline/db_utils/heavy_ranker.py) on the B200 engine.

Same flow, same names: build ``documents.db`` with the corpus builder (setup_docs_db.py), index the
passages twice with ``hybrid=True, content=True`` (:78-89), load the two indexes (:91-94), search each
query with limit 1 (:98-101), fetch the passages from sqlite (:102-109) and apply the agreement rule
(:110-115) -- once the reference's way (one query at a time through ``txtai``-named calls) and once
batched through ``HeavyRanker``.

The sentence-transformers weights and the Wikipedia dump are not available offline, so a deterministic
stand-in encoder (text -> seeded Gaussian, a real encoder goes in through ``path=`` / ``transform=``)
and a few synthetic articles are used.  Needs a B200:  python examples/heavy_ranker_b200.py
"""
import os
import sys
import tempfile
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import vietnamese_qa_system_b200 as txtai  # noqa: E402  (heavy_ranker.py:4  import txtai)
from vietnamese_qa_system_b200 import HeavyRanker, corpus  # noqa: E402
from vietnamese_qa_system_b200.db import query  # noqa: E402  (heavy_ranker.py:6-7  from setup_db import ...)


class StandInEncoder:
    """text -> vector; stands in for paraphrase-multilingual-{MiniLM-L12-v2 (384), mpnet-base-v2 (768)}."""

    def __init__(self, dim):
        self.dim = dim

    def __call__(self, texts):
        out = np.empty((len(texts), self.dim), np.float32)
        for r, t in enumerate(texts):
            out[r] = np.random.default_rng(zlib.crc32(t.encode("utf-8"))).standard_normal(self.dim)
        return out


ARTICLES = [
    "Hà_Nội là thủ_đô của Việt_Nam. Thành_phố nằm bên sông Hồng. " * 12,
    "Phở là món ăn nổi_tiếng của ẩm_thực Việt_Nam, gồm bánh phở, nước dùng và thịt bò. " * 10,
    "Vịnh Hạ_Long ở tỉnh Quảng_Ninh là di_sản thiên_nhiên thế_giới với hàng nghìn đảo đá vôi. " * 9,
    "Trí_tuệ nhân_tạo là lĩnh_vực của khoa_học máy_tính nghiên_cứu các hệ_thống thông_minh. " * 11,
]


def main():
    work = tempfile.mkdtemp(prefix="vqa_demo_")
    database = os.path.join(work, "documents.db")
    passages = corpus.insert_doc(database, texts=ARTICLES)                      # setup_docs_db.py:16-52
    print(f"{len(passages)} passages in {database}")

    # ---- build + save (:70-89) ---------------------------------------------------------------------
    data = query(database, query_string="SELECT * FROM documents", fetch_size=50000)
    data_str = [{"id": row[0], "text": row[1], "source": row[2]} for row in data]
    enc = {"mini_lm": StandInEncoder(384), "mpnet": StandInEncoder(768)}
    for name in ("mini_lm", "mpnet"):
        embeddings = txtai.Embeddings(hybrid=True, content=True, transform=enc[name])
        embeddings.index(data_str)
        embeddings.save(os.path.join(work, "embeddings_index", name))

    # ---- load (:91-94) ---------------------------------------------------------------------------------
    embeddings_MiniLM = txtai.Embeddings(transform=enc["mini_lm"])
    embeddings_MiniLM.load(os.path.join(work, "embeddings_index", "mini_lm"))
    embeddings_mpnet = txtai.Embeddings(transform=enc["mpnet"])
    embeddings_mpnet.load(os.path.join(work, "embeddings_index", "mpnet"))

    sample_queries = (passages[0], "phở bò Hà Nội", "di sản thiên nhiên thế giới ở Quảng Ninh",
                      "How do electric cars work?")
    # ---- the reference's loop (:97-115) -------------------------------------------------------------
    for query_str in sample_queries:
        semantic_MiniLM = embeddings_MiniLM.search(query_str, 1)[0]
        uid_paraphrase, score_paraphrase = semantic_MiniLM["id"], semantic_MiniLM["score"]
        semantic_mpnet = embeddings_mpnet.search(query_str, 1)[0]
        uid_mpnet, score_mpnet = semantic_mpnet["id"], semantic_mpnet["score"]
        doc = query(database, query_string=f"SELECT doc FROM documents WHERE id = {uid_mpnet}", fetch_size=1)
        if uid_paraphrase == uid_mpnet and score_paraphrase + score_mpnet > 0.4:
            print(f"MATCH  {query_str[:40]!r:44} id {uid_mpnet}  score {score_paraphrase + score_mpnet:.4f}  {doc[0][:50]!r}")
        else:
            print(f"no doc {query_str[:40]!r:44} ids {uid_paraphrase}/{uid_mpnet}  "
                  f"scores {score_paraphrase:.4f}/{score_mpnet:.4f}")

    # ---- the same, batched: two searches, one agreement kernel, one sqlite fetch -------------------------
    ranker = HeavyRanker(embeddings_MiniLM, embeddings_mpnet, database_path=database)
    for q, r in zip(sample_queries, ranker.rank(list(sample_queries))):
        print(f"batched {q[:30]!r:34} match={r['match']}  context={ranker.context(r)[:60]!r}")


if __name__ == "__main__":
    main()
