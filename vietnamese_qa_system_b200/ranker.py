"""The reference's retriever call site, batched: ``heavy_ranker.py:97-115``.

The reference loops its queries one at a time, asks two indexes (MiniLM-L12,
D=384 and mpnet-base, D=768) for their top-1, fetches both passages from sqlite
and accepts a passage when both indexes agree and the two scores sum to more than
0.4 (:110).  ``HeavyRanker`` does the same for a whole batch: one batched search
per index, one batched agreement kernel (``vqa_agree``), one batched sqlite fetch.
``straighten_docs`` renders accepted passages the way the QA prompt consumes them
(src/data/configs/advance_qa_sample.py:99-106).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import torch

import os

from . import ops
from .db import fetch_docs, query
from .embeddings import Embeddings

# heavy_ranker.py:78-83 -- the two encoders and the configuration the reference builds its indexes with
REFERENCE_INDEXES = {
    "mini_lm": {"hybrid": True, "content": True,
                "path": "sentence-transformers/paraphrase-multilingual-MiniLM-L12-v2"},
    "mpnet": {"hybrid": True, "content": True,
              "path": "sentence-transformers/paraphrase-multilingual-mpnet-base-v2"},
}


def load_passages(database_path: str, fetch_size: int = 50000) -> List[Dict[str, Any]]:
    """``documents.db`` rows as the dicts the reference indexes (heavy_ranker.py:70-76)."""
    data = query(database_path, query_string="SELECT * FROM documents", fetch_size=fetch_size)
    return [{"id": row[0], "text": row[1], "source": row[2]} for row in data]

# src/data/configs/response_template.py:285 (NO_DOCS_MESSAGE1, what get_no_docs_msg(id=1) returns)
NO_DOCS_MESSAGE = " Không documents nào có điểm đủ cao để query cho câu hỏi. "


def straighten_docs(docs_list: Sequence[str]) -> str:
    """The [CONTEXT] payload of the QA prompt, exactly as advance_qa_sample.py:99-106 renders it:
    ``" [CTX{i}]: {doc} [ECTX{i}] "`` joined over passages (0-based), or the bracketed
    no-documents message when the list is empty."""
    if not docs_list:
        return f"[ERROR]{NO_DOCS_MESSAGE}[ERROR]"
    return "".join(f" [CTX{idx}]: {doc} [ECTX{idx}] " for idx, doc in enumerate(docs_list))


class HeavyRanker:
    def __init__(self, embeddings_a: Embeddings, embeddings_b: Embeddings, database_path: Optional[str] = None,
                 threshold: float = 0.4):
        self.a, self.b = embeddings_a, embeddings_b
        self.database_path = database_path
        self.threshold = float(threshold)

    @classmethod
    def build(cls, database_path: str, index_dir: str, configs: Optional[Dict[str, dict]] = None,
              embeddings_cls=Embeddings, threshold: float = 0.4, **overrides) -> "HeavyRanker":
        """The build half of heavy_ranker.py (:70-89): read the passages, index them with each of the two
        configurations, save under ``index_dir/<name>``.  ``overrides`` (e.g. ``transform=``, ``dtype=``) are
        merged into both configurations."""
        configs = configs or REFERENCE_INDEXES
        if len(configs) != 2:
            raise ValueError("the agreement rule (heavy_ranker.py:110) is defined over exactly two indexes")
        data_str = load_passages(database_path)
        built = []
        for name, cfg in configs.items():
            per = {k: (v[name] if isinstance(v, dict) and set(v) == set(configs) else v) for k, v in overrides.items()}
            emb = embeddings_cls(**{**cfg, **per})
            emb.index(data_str)
            emb.save(os.path.join(index_dir, name))
            built.append(emb)
        return cls(built[0], built[1], database_path=database_path, threshold=threshold)

    @classmethod
    def load(cls, database_path: str, index_dir: str, names=("mini_lm", "mpnet"), embeddings_cls=Embeddings,
             threshold: float = 0.4, **overrides) -> "HeavyRanker":
        """The load half (heavy_ranker.py:91-94)."""
        loaded = []
        for name in names:
            per = {k: (v[name] if isinstance(v, dict) and set(v) == set(names) else v) for k, v in overrides.items()}
            emb = embeddings_cls(**per)
            emb.load(os.path.join(index_dir, name))
            loaded.append(emb)
        return cls(loaded[0], loaded[1], database_path=database_path, threshold=threshold)

    def rank(self, queries_a: Any, queries_b: Any = None, limit: int = 1) -> List[Dict[str, Any]]:
        """``queries_a`` / ``queries_b``: the same queries as each index's encoder sees them (texts, or
        per-index vectors since the two indexes have different dimensions)."""
        queries_b = queries_a if queries_b is None else queries_b
        as_list = lambda q: list(q) if isinstance(q, (list, tuple)) else q  # noqa: E731
        sa, pa = self.a.search_tensors(as_list(queries_a), limit)   # float32 (dense) or float64 (hybrid) scores
        sb, pb = self.b.search_tensors(as_list(queries_b), limit)
        if sa.dtype != sb.dtype:                                   # one dense, one hybrid index
            sa, sb = sa.double(), sb.double()
        # map ANN positions to caller ids before comparing (heavy_ranker.py:99,101 compare r['id'])
        ida = self._user_ids(self.a, pa[:, 0])
        idb = self._user_ids(self.b, pb[:, 0])
        accept, combined = ops.agree(ida, sa[:, 0].contiguous(), idb, sb[:, 0].contiguous(), self.threshold)
        ida_h, idb_h = ida.cpu().tolist(), idb.cpu().tolist()
        sa_h, sb_h = sa[:, 0].cpu().tolist(), sb[:, 0].cpu().tolist()
        acc_h, comb_h = accept.cpu().tolist(), combined.cpu().tolist()
        texts: Dict[int, str] = {}
        if self.database_path is not None:
            texts = fetch_docs(self.database_path, ida_h + idb_h)
        out = []
        for i in range(len(ida_h)):
            out.append({"id_a": ida_h[i], "score_a": sa_h[i], "id_b": idb_h[i], "score_b": sb_h[i],
                        "match": bool(acc_h[i]), "score": comb_h[i] if acc_h[i] else None,
                        "doc_a": texts.get(ida_h[i]), "doc_b": texts.get(idb_h[i])})
        return out

    @staticmethod
    def _user_ids(emb: Embeddings, pos: torch.Tensor) -> torch.Tensor:
        if emb._id_is_position:
            return pos.contiguous()
        table = torch.as_tensor([int(u) for u in emb.ids], dtype=torch.int64, device=pos.device)
        return table[pos.clamp_min(0)].contiguous()

    def context(self, result: Dict[str, Any]) -> str:
        """Prompt context for one ranked query: the agreed passage, else the no-doc message."""
        if result["match"] and result.get("doc_a"):
            return straighten_docs([result["doc_a"]])
        return straighten_docs([])
