"""``Embeddings`` -- drop-in for the reference's retriever object.

The reference builds and queries its document index through ``txtai.Embeddings``
(inference_pipeline/db_utils/heavy_ranker.py:78-101):

    embeddings = txtai.Embeddings(hybrid=True, content=True, path="sentence-transformers/...")
    embeddings.index([{"id": ..., "text": ..., "source": ...}, ...])     # :86,88
    embeddings.save("./inference_pipeline/embeddings_index/mpnet")         # :87,89
    embeddings = txtai.Embeddings(); embeddings.load(<dir>)                # :91-94
    hit = embeddings.search(query_str, 1)[0]; hit['id'], hit['score']      # :98-101

This class keeps those names, positional orders, defaults (``limit=3``) and result
shapes (``[{"id","text","score"}]`` with ``content=True``, else ``[(id, score)]``) so
that ``import vietnamese_qa_system_b200 as txtai`` leaves heavy_ranker.py unchanged.
The dense leg (encode -> pool -> normalise -> score -> top-k) runs on the GPU
through ``libvqa_b200.so``.  ``hybrid=True`` (how the reference builds its indexes,
:78,81) adds the sparse BM25 leg (``scoring.BM25``): both legs
fetch ``10 * limit`` candidates on the GPU and ``vqa_hybrid_fuse`` adds their scores per
id with weights ``[w, 1 - w]`` (w = 0.5), as txtai does.

Vectors: queries / documents may be given as text (needs an encoder: ``transform=``
callable, or a HuggingFace ``path`` loaded by ``vectors.HFEncoder``) or directly as
float arrays ``[D]`` / ``[B, D]`` (txtai's ``method="external"`` convention), which is
what every BASELINE.json configuration uses.
"""
from __future__ import annotations

import json
import os
import sqlite3
from typing import Any, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .ann import B200Flat
from .scoring import BM25

_CONFIG_FILE = "config.json"
_ANN_FILE = "embeddings"
_DB_FILE = "documents"
_IDS_FILE = "ids.json"
_SCORING_FILE = "scoring"
_HYBRID_CANDIDATES = 10  # each leg fetches limit * 10 candidates (txtai Search)
_DENSE_K_MAX = 1024      # vqa_search answers k <= 128 per call; ops.FlatShard composes up to 1024 (vqa_merge_segments)
_PERSISTED_KEYS_SKIP = {"transform", "device", "shards", "exchange"}   # deployment choices, not index properties


def _is_vector(x: Any) -> bool:
    return isinstance(x, (np.ndarray, torch.Tensor)) or (
        isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], (float, int, np.floating)))


class Embeddings:
    def __init__(self, config: Optional[dict] = None, **kwargs):
        self.config: dict = {}
        self.ann: Optional[B200Flat] = None
        self.ids: List[Any] = []          # ANN position -> caller's id (heavy_ranker.py:74-76)
        self._id_is_position = True
        self.database: Optional[sqlite3.Connection] = None
        self.scoring: Optional[BM25] = None  # sparse leg (hybrid=True / keyword=True / scoring={...})
        self._encoder = None
        self.configure({**(config or {}), **kwargs})

    # ------------------------------------------------------------------ config
    def configure(self, config: dict) -> None:
        self.config = dict(config)
        if self.config.get("hybrid") or self.config.get("keyword"):
            # txtai: hybrid=True -> dense index + BM25 term index with normalised scores;
            # keyword=True -> the term index alone
            self.config.setdefault("scoring", {"method": "bm25", "normalize": bool(self.config.get("hybrid")),
                                               "terms": True})
        sc = self.config.get("scoring")
        if isinstance(sc, str):
            self.config["scoring"] = sc = {"method": sc, "terms": True}
        if sc is not None and str(sc.get("method", "bm25")).lower() != "bm25":
            raise NotImplementedError(f"scoring method {sc.get('method')!r}: only bm25 is built")
        self._transform = self.config.get("transform")

    @property
    def _dense(self) -> bool:
        return not self.config.get("keyword")

    @property
    def content(self) -> bool:
        return bool(self.config.get("content"))

    def _ann_config(self) -> dict:
        cfg = {"dtype": self.config.get("dtype", "bf16"), "device": self.config.get("device")}
        for key in ("mode", "exchange"):
            if key in self.config:
                cfg[key] = self.config[key]
        return cfg

    def _new_ann(self) -> B200Flat:
        """``shards=True`` (under torchrun, one process per GPU): the dense index is row-sharded over the ranks and
        searched with one NVLink exchange per batch (``ann.B200Sharded``); every rank makes the same calls and gets
        the same results.  The BM25 leg of a hybrid index and the content store are replicated per rank -- they are
        small next to the dense matrix.  Default: one GPU holds the whole index (``ann.B200Flat``)."""
        if self.config.get("shards"):
            from .ann import B200Sharded

            return B200Sharded(self._ann_config())
        return B200Flat(self._ann_config())

    # ------------------------------------------------------------------ vectors
    def _encoder_fn(self):
        if self._transform is not None:
            return self._transform
        if self._encoder is None:
            path = self.config.get("path")
            if not path:
                raise ValueError("text input needs an encoder: pass transform=callable or path=<HF model dir>")
            from .vectors import HFEncoder

            # encoderdtype: precision of the HF forward.  Default fp32 = what the reference runs (txtai never
            # down-casts the encoder); "bf16"/"fp16" trade ~1e-2 relative embedding error for speed.  The key is
            # part of the config, so it is saved with the index and a reloaded index encodes queries the same way.
            self._encoder = HFEncoder(path, device=self.config.get("device"),
                                      batch=int(self.config.get("encodebatch", 32)),
                                      maxlength=self.config.get("maxlength"),
                                      dtype=self.config.get("encoderdtype", "fp32"))
        return self._encoder

    def _device(self) -> torch.device:
        from . import _native

        _native.require_cuda()
        d = self.config.get("device")
        return torch.device("cuda", torch.cuda.current_device()) if d is None else torch.device(d)

    def batchtransform(self, documents: Sequence[Any]) -> torch.Tensor:
        """Texts or vectors -> L2-normalised float32 CUDA tensor [B, D]."""
        dev = self._device()
        if isinstance(documents, torch.Tensor):
            t = documents
        elif isinstance(documents, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(documents, dtype=np.float32))
        else:
            data = [d[1] if isinstance(d, tuple) and len(d) == 3 else d for d in documents]
            if len(data) == 0:
                raise ValueError("no documents / queries given")
            if all(isinstance(d, str) for d in data):
                enc = self._encoder_fn()
                out = enc(data)
                t = out if isinstance(out, torch.Tensor) else torch.from_numpy(np.asarray(out, dtype=np.float32))
                if t.is_cuda and t.dtype == torch.float32 and getattr(enc, "normalized", False):
                    return t  # K1 already pooled + normalised on device
            elif all(_is_vector(d) for d in data):
                t = torch.stack([d.detach().to("cpu", torch.float32) if isinstance(d, torch.Tensor)
                                 else torch.from_numpy(np.asarray(d, dtype=np.float32)) for d in data])
            else:
                raise ValueError("documents / queries must be all text or all vectors")
        t = t.to(device=dev, dtype=torch.float32)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if t.dim() != 2:
            raise ValueError(f"expected [B, D] vectors; got shape {tuple(t.shape)}")
        return ops.normalize_rows(t.contiguous())

    def transform(self, document: Any) -> torch.Tensor:
        return self.batchtransform([document])[0]

    # ------------------------------------------------------------------ build
    @staticmethod
    def _unpack(doc: Any, auto_id: int) -> Tuple[Any, Any, Any]:
        """txtai document forms: (id, data, tags) | dict with id/text | bare str/vector."""
        if isinstance(doc, tuple) and len(doc) == 3:
            return doc
        if isinstance(doc, tuple) and len(doc) == 2:
            return doc[0], doc[1], None
        if isinstance(doc, dict):
            uid = doc.get("id", auto_id)
            return uid, doc, None
        return auto_id, doc, None

    def _open_db(self, path: Optional[str] = None) -> sqlite3.Connection:
        con = sqlite3.connect(path or ":memory:")
        con.execute("CREATE TABLE IF NOT EXISTS sections (indexid INTEGER PRIMARY KEY, id TEXT, text TEXT, data TEXT)")
        return con

    def index(self, documents: Iterable[Any], embeddings=None) -> None:
        """Build the index.  ``documents``: iterable of dict{id,text,...} (the reference's form),
        (id, data, tags) tuples, strings, or vectors.  ``embeddings`` (optional, [N,D]) supplies
        pre-computed vectors for text documents (config A-D path)."""
        self.ids, rows = [], []
        for n, doc in enumerate(documents):
            uid, data, _ = self._unpack(doc, n)
            self.ids.append(uid)
            rows.append(data)
        self._id_is_position = all(isinstance(u, int) and u == p for p, u in enumerate(self.ids))
        payload = [r.get("text") if isinstance(r, dict) else r for r in rows]
        self.scoring = None
        if self.config.get("scoring"):
            if not all(isinstance(t, str) for t in payload):
                raise ValueError("hybrid / keyword indexes need text documents (the sparse leg scores terms)")
            self.scoring = BM25(self.config["scoring"], device=self.config.get("device"))
            self.scoring.index(payload)
        if not self._dense:
            self.ann = None
            self._store_content(rows)
            return
        if embeddings is not None:
            vecs = self.batchtransform(embeddings)
            if vecs.shape[0] != len(rows):
                raise ValueError(f"{len(rows)} documents but {vecs.shape[0]} embeddings")
        else:
            batch = int(self.config.get("batch", 500))
            parts = [self.batchtransform(payload[i:i + batch]) for i in range(0, len(payload), batch)]
            vecs = torch.cat(parts) if parts else torch.empty((0, int(self.config.get("dimensions", 0))),
                                                              device="cuda")
        self.config["dimensions"] = int(vecs.shape[1])
        self.ann = self._new_ann()
        self.ann.index(vecs)
        self._store_content(rows)

    def _store_content(self, rows: List[Any]) -> None:
        if self.content:
            self.database = self._open_db()
            recs = []
            for pos, (uid, r) in enumerate(zip(self.ids, rows)):
                text = r.get("text") if isinstance(r, dict) else (r if isinstance(r, str) else None)
                extra = json.dumps({k: v for k, v in r.items() if k not in ("id", "text")}, ensure_ascii=False) \
                    if isinstance(r, dict) else None
                recs.append((pos, str(uid), text, extra))
            with self.database:
                self.database.executemany("INSERT INTO sections VALUES (?, ?, ?, ?)", recs)

    def count(self) -> int:
        if self.ann is not None:
            return self.ann.count()
        return 0 if self.scoring is None else self.scoring.count()

    # ------------------------------------------------------------------ query
    def search(self, query: Any, limit: Optional[int] = None, weights=None, index=None, parameters=None,
               graph: bool = False):
        """Top-``limit`` (default 3) results for one query string / vector."""
        return self.batchsearch([query], limit, weights, index, parameters, graph)[0]

    def batchsearch(self, queries: Sequence[Any], limit: Optional[int] = None, weights=None, index=None,
                    parameters=None, graph: bool = False):
        if graph:
            raise NotImplementedError("graph search is not part of the retrieval hot path")
        limit = 3 if limit is None else int(limit)
        scores, pos = self.search_tensors(queries, limit, weights)
        s_np, p_np = scores.cpu().numpy(), pos.cpu().numpy()
        results = []
        for b in range(s_np.shape[0]):
            hits = [(int(p), float(s)) for p, s in zip(p_np[b].tolist(), s_np[b].tolist()) if p >= 0]
            results.append(self._resolve(hits))
        return results

    def search_tensors(self, queries: Sequence[Any], limit: int, weights=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-resident search: ``(scores [B, k], positions int64 [B, k])`` with ``k = min(limit, count)``;
        scores are float32 cosines for a dense index and float64 for hybrid / keyword indexes (the Python
        floats txtai returns); unused slots hold ``-inf`` / ``-1``."""
        if self.ann is None and self.scoring is None:
            raise RuntimeError("index is empty: call index() or load() first")
        limit = int(limit)
        if limit < 1:
            raise ValueError(f"limit must be >= 1; got {limit}")
        queries = queries if isinstance(queries, (np.ndarray, torch.Tensor)) else list(queries)
        k = min(limit, max(self.count(), 1))
        if self.scoring is None:
            return self.ann.search_tensors(self.batchtransform(queries), k)
        if isinstance(queries, (np.ndarray, torch.Tensor)) or not all(isinstance(q, str) for q in queries):
            raise ValueError("hybrid / keyword search needs text queries (the sparse leg scores terms)")
        if self.ann is None:
            return self.scoring.search_tensors(queries, k)
        cand = limit * _HYBRID_CANDIDATES
        kd = min(cand, self.count())
        if kd > _DENSE_K_MAX:
            raise NotImplementedError(f"hybrid search fetches {_HYBRID_CANDIDATES} x limit dense candidates and the "
                                      f"engine answers k <= {_DENSE_K_MAX}: limit <= "
                                      f"{_DENSE_K_MAX // _HYBRID_CANDIDATES} is supported (got {limit})")
        if weights is None:
            weights = 0.5
        if isinstance(weights, (int, float)):
            weights = [weights, 1 - weights]
        ds, dp = self.ann.search_tensors(self.batchtransform(queries), kd)
        ss, sp = self.scoring.search_tensors(queries, cand)
        # txtai: weighted score sum only when the sparse scores are normalised to [0, 1]; raw BM25 scores are
        # unbounded, so an un-normalised scoring index is fused by reciprocal rank.  Legs with weight <= 0 are
        # ignored by the kernel (weights 1 / 0 = the single-leg answer, as in txtai).
        rrf = not bool(getattr(self.scoring, "normalize", True))
        return ops.hybrid_fuse(ds, dp, ss, sp, min(limit, kd + cand), float(weights[0]), float(weights[1]), rrf=rrf)

    def _resolve(self, hits: List[Tuple[int, float]]):
        """ANN position -> caller's id (a6), and the content join when content=True."""
        if self.content and self.database is not None:
            if not hits:
                return []
            marks = ",".join("?" * len(hits))
            rows = self.database.execute(f"SELECT indexid, id, text FROM sections WHERE indexid IN ({marks})",
                                         [p for p, _ in hits]).fetchall()
            by_pos = {r[0]: r for r in rows}
            out = []
            for p, s in hits:
                r = by_pos.get(p)
                uid = self.ids[p] if p < len(self.ids) else (r[1] if r else p)
                out.append({"id": uid, "text": r[2] if r else None, "score": s})
            return out
        return [((self.ids[p] if p < len(self.ids) else p), s) for p, s in hits]

    def similarity(self, query: Any, data: Sequence[Any]) -> List[Tuple[int, float]]:
        """Score ``query`` against ad-hoc ``data`` (txtai API): [(index, score)] descending."""
        tmp = Embeddings({**{k: v for k, v in self.config.items() if k not in ("scoring", "keyword")},
                          "content": False, "hybrid": False}, transform=self._transform)
        tmp._encoder = self._encoder
        tmp.index(list(data))
        return tmp.search(query, len(data))

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        """Directory with config + embeddings (+ documents when content=True)."""
        if self.ann is None and self.scoring is None:
            raise RuntimeError("nothing to save")
        os.makedirs(path, exist_ok=True)
        cfg = {k: v for k, v in self.config.items() if k not in _PERSISTED_KEYS_SKIP and _jsonable(v)}
        with open(os.path.join(path, _CONFIG_FILE), "w", encoding="utf-8") as f:
            json.dump(cfg, f, ensure_ascii=False)
        if self.ann is not None:
            self.ann.save(os.path.join(path, _ANN_FILE))
        if self.scoring is not None:
            self.scoring.save(os.path.join(path, _SCORING_FILE))
        with open(os.path.join(path, _IDS_FILE), "w", encoding="utf-8") as f:
            json.dump(None if self._id_is_position else self.ids, f, ensure_ascii=False)
        if self.content and self.database is not None:
            target = os.path.join(path, _DB_FILE)
            if os.path.exists(target):
                os.remove(target)
            disk = sqlite3.connect(target)
            with disk:
                self.database.backup(disk)
            disk.close()

    def load(self, path: str) -> "Embeddings":
        with open(os.path.join(path, _CONFIG_FILE), "r", encoding="utf-8") as f:
            cfg = json.load(f)
        keep = {k: v for k, v in self.config.items() if k in _PERSISTED_KEYS_SKIP}
        self.configure({**cfg, **keep})
        self.ann = self.scoring = None
        if self._dense:
            self.ann = self._new_ann()
            self.ann.load(os.path.join(path, _ANN_FILE))
        if self.config.get("scoring"):
            self.scoring = BM25(self.config["scoring"], device=self.config.get("device"))
            self.scoring.load(os.path.join(path, _SCORING_FILE))
        with open(os.path.join(path, _IDS_FILE), "r", encoding="utf-8") as f:
            ids = json.load(f)
        self._id_is_position = ids is None
        self.ids = list(range(self.count())) if ids is None else ids
        db = os.path.join(path, _DB_FILE)
        self.database = None
        if self.content and os.path.exists(db):
            disk = sqlite3.connect(db)
            self.database = sqlite3.connect(":memory:")
            disk.backup(self.database)
            disk.close()
        return self

    def exists(self, path: str) -> bool:
        return os.path.exists(os.path.join(path, _CONFIG_FILE))

    def close(self) -> None:
        if self.database is not None:
            self.database.close()
            self.database = None
        self.ann = self.scoring = None


def _jsonable(v: Any) -> bool:
    try:
        json.dumps(v)
        return True
    except (TypeError, ValueError):
        return False
