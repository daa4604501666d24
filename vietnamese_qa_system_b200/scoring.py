"""Sparse keyword leg of hybrid search: tokeniser + BM25 term index resident in HBM.

The reference's indexes are built ``txtai.Embeddings(hybrid=True, content=True, path=...)``
(inference_pipeline/db_utils/heavy_ranker.py:78-83).  Behind ``hybrid=True`` txtai configures
``scoring = {"method": "bm25", "terms": True, "normalize": True}``: a BM25 term index over the
same document positions as the dense index, queried for ``10 * limit`` candidates whose scores
are then added to the dense leg's.  This module is that
scoring object: same configuration keys (``k1``, ``b``, ``normalize``, ``terms.cutoff``,
``tokenizer``), same ``index / search / batchsearch / count / save / load`` surface.

Host side (this file): tokenising text and laying the postings out as CSR -- string work, done
once at build time.  Device side (``libvqa_b200.so``): the BM25 weight of every posting
(``vqa_bm25_weights``), and per query batch the accumulate + top-k (``vqa_sparse_search``).
There is no CPU scoring path.

Two orders txtai leaves to chance are pinned here: equal scores come back lower position
first (txtai: ``argpartition``/unstable ``argsort``), and document-frequency statistics are
taken in first-occurrence order of the terms (txtai iterates a ``set``; this only affects the
last bit of ``avgidf``).
"""
from __future__ import annotations

import ctypes
import json
import os
import string
from collections import Counter
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as N

_META_STRIDE = 4


class Tokenizer:
    """txtai's ``Tokenizer``: the default (``alphanum=False``) is Unicode word segmentation
    (UAX #29) of the lower-cased text, which keeps Vietnamese syllables with their diacritics."""

    STOP_WORDS = {"a", "an", "and", "are", "as", "at", "be", "but", "by", "for", "if", "in", "into", "is", "it",
                  "no", "not", "of", "on", "or", "such", "that", "the", "their", "then", "there", "these", "they",
                  "this", "to", "was", "will", "with"}

    def __init__(self, lowercase: bool = True, emoji: bool = True, alphanum: bool = False,
                 stopwords: Union[bool, Sequence[str]] = False):
        self.lowercase = lowercase
        self.alphanum = self.segment = None
        if alphanum:
            import re

            # tokens of >= 2 characters with at least one non-trailing letter
            self.alphanum = re.compile(r"^\d*[a-z][\-.0-9:_a-z]{1,}$")
        else:
            import regex

            pattern = r"\w\p{Extended_Pictographic}\p{WB:RegionalIndicator}" if emoji else r"\w"
            self.segment = regex.compile(rf"[{pattern}](?:\B\S)*", flags=regex.WORD)
        self.stopwords = set(stopwords) if isinstance(stopwords, (list, tuple, set)) else \
            (Tokenizer.STOP_WORDS if stopwords else None)

    def __call__(self, text: str) -> List[str]:
        text = text.lower() if self.lowercase else text
        if self.alphanum is not None:
            tokens = [t.strip(string.punctuation) for t in text.split()]
            tokens = [t for t in tokens if self.alphanum.match(t)]
        else:
            tokens = self.segment.findall(text)
        if self.stopwords:
            tokens = [t for t in tokens if t not in self.stopwords]
        return tokens


class BM25:
    """BM25 term index over document positions ``0..N-1``; postings and weights live on the GPU."""

    def __init__(self, config: Optional[dict] = None, device=None):
        self.config = dict(config or {})
        self.k1 = float(self.config.get("k1", 1.2))
        self.b = float(self.config.get("b", 0.75))
        self.normalize = bool(self.config.get("normalize", False))
        terms = self.config.get("terms")
        self.cutoff = float(terms.get("cutoff", 0.1)) if isinstance(terms, dict) else 0.1
        tok = self.config.get("tokenizer")
        self.tokenizer = Tokenizer(**tok) if isinstance(tok, dict) else (tok if callable(tok) else Tokenizer())
        self._device_arg = device
        self.vocab: Dict[str, int] = {}
        self.total = 0
        self.tokens = 0
        self.avgdl = self.avgfreq = self.avgidf = 0.0
        self.avgscore: Optional[float] = None
        self._h = None
        self._ws: Dict[Tuple[int, int], torch.Tensor] = {}
        self._host: Dict[str, np.ndarray] = {}
        self._dev: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ helpers
    @property
    def device(self) -> torch.device:
        N.require_cuda()
        d = self._device_arg
        if d is None:
            return torch.device("cuda", torch.cuda.current_device())
        d = torch.device(d)
        return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())

    def tokenize(self, text: Any) -> List[str]:
        return self.tokenizer(text) if isinstance(text, str) else list(text)

    def count(self) -> int:
        return self.total

    def __del__(self):
        self._release()

    def _release(self) -> None:
        h = getattr(self, "_h", None)
        lib = getattr(N, "_lib", None) if N is not None else None
        if h is not None and lib is not None:
            lib.vqa_sparse_destroy(h)
        self._h = None

    # ------------------------------------------------------------------ build
    def index(self, documents: Iterable[Any]) -> None:
        """``documents``: texts (or token lists) in document-position order."""
        self.build_postings(documents)
        self._upload()

    def build_postings(self, documents: Iterable[Any]) -> None:
        """Host half of ``index``: tokenise, CSR postings, scalar statistics (no device work)."""
        vocab: Dict[str, int] = {}
        t_ids: List[int] = []
        t_docs: List[int] = []
        t_freqs: List[int] = []
        lengths: List[int] = []
        wordfreq: List[int] = []
        for pos, doc in enumerate(documents):
            if isinstance(doc, dict):
                doc = doc.get("text")
            tokens = self.tokenize(doc) if doc is not None else []
            lengths.append(len(tokens))
            for term, freq in Counter(tokens).items():  # first-occurrence order
                tid = vocab.get(term)
                if tid is None:
                    tid = vocab[term] = len(vocab)
                    wordfreq.append(0)
                wordfreq[tid] += freq
                t_ids.append(tid)
                t_docs.append(pos)
                t_freqs.append(freq)
        if len(lengths) >= 2 ** 31:
            raise ValueError("at most 2^31 - 1 documents per term index")
        self.vocab = vocab
        self.total = len(lengths)
        tid_a = np.asarray(t_ids, dtype=np.int64)
        order = np.argsort(tid_a, kind="stable")  # postings of a term stay in ascending position order
        n_terms = len(vocab)
        df = np.bincount(tid_a, minlength=n_terms).astype(np.int64) if n_terms else np.zeros(0, np.int64)
        offsets = np.zeros(n_terms + 1, dtype=np.int64)
        np.cumsum(df, out=offsets[1:])
        self._host = {
            "offsets": offsets,
            "docs": np.asarray(t_docs, dtype=np.int32)[order],
            "freqs": np.asarray(t_freqs, dtype=np.int32)[order],
            "lengths": np.asarray(lengths, dtype=np.int32),
            "wordfreq": np.asarray(wordfreq, dtype=np.int64),
        }
        self._stats()

    def index_postings(self, offsets, docs, freqs, lengths, vocab: Optional[Sequence[str]] = None) -> None:
        """Build from ready-made CSR postings (an external tokeniser / analyser): ``offsets`` int64[T+1],
        ``docs`` / ``freqs`` int32[P] (positions ascending within a term), ``lengths`` int32[N] tokens per
        document.  ``vocab``: the T term strings (default: ``str(term_id)``)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        docs = np.ascontiguousarray(docs, dtype=np.int32)
        freqs = np.ascontiguousarray(freqs, dtype=np.int32)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        n_terms = len(offsets) - 1
        if n_terms < 0 or offsets[0] != 0 or offsets[-1] != len(docs) or len(docs) != len(freqs) \
                or np.any(np.diff(offsets) < 1):
            raise ValueError("offsets must start at 0, end at len(docs) and give every term >= 1 posting")
        wordfreq = np.add.reduceat(freqs.astype(np.int64), offsets[:-1]) if n_terms else np.zeros(0, np.int64)
        names = [str(t) for t in range(n_terms)] if vocab is None else list(vocab)
        if len(names) != n_terms:
            raise ValueError(f"{n_terms} terms but {len(names)} vocabulary entries")
        self.vocab = {t: i for i, t in enumerate(names)}
        self.total = int(len(lengths))
        self._host = {"offsets": offsets, "docs": docs, "freqs": freqs, "lengths": lengths, "wordfreq": wordfreq}
        self._stats()
        self._upload()

    def _stats(self) -> None:
        """Scalar statistics (TFIDF.index): token counts, idf per term, avgidf, avgscore."""
        h = self._host
        df = np.diff(h["offsets"])
        n_terms = len(df)
        self.tokens = int(h["wordfreq"].sum())
        self.idf_host = np.zeros(0, dtype=np.float64)
        self.avgscore = None
        if n_terms:
            self.avgfreq = self.tokens / n_terms
            self.avgdl = self.tokens / self.total
            self.idf_host = np.log(1 + (self.total - df + 0.5) / (df + 0.5))
            self.avgidf = float(np.mean(self.idf_host))
            k = self.k1 * ((1 - self.b) + self.b * self.avgdl / self.avgdl)
            self.avgscore = float(self.avgidf * (self.avgfreq * (self.k1 + 1)) / (self.avgfreq + k))
        self._df = df
        self._common = None

    def _upload(self) -> None:
        """Device half of ``index`` / ``load``: postings to HBM, BM25 weight per posting (``vqa_bm25_weights``)."""
        h = self._host
        n_terms = len(self._df)
        dev = self.device
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        d = {"offsets": up(h["offsets"]),
             "docs": up(h["docs"]) if len(h["docs"]) else torch.zeros(1, dtype=torch.int32, device=dev)}
        n_post = int(len(h["docs"]))
        weights = torch.zeros(max(n_post, 1), dtype=torch.float32, device=dev)
        if n_post:
            freqs, lens, idf = up(h["freqs"]), up(h["lengths"]), up(self.idf_host)
            N.check(N.lib().vqa_bm25_weights(
                ctypes.c_void_p(d["offsets"].data_ptr()), n_terms, ctypes.c_void_p(d["docs"].data_ptr()),
                ctypes.c_void_p(freqs.data_ptr()), n_post, ctypes.c_void_p(idf.data_ptr()),
                ctypes.c_void_p(lens.data_ptr()), self.k1, self.b, float(self.avgdl),
                ctypes.c_void_p(weights.data_ptr()), dev.index or 0,
                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            torch.cuda.current_stream(dev).synchronize()  # freqs / lens / idf are freed on return
        d["weights"] = weights
        self._dev = d
        self._release()
        hnd = ctypes.c_void_p()
        N.check(N.lib().vqa_sparse_create(ctypes.byref(hnd), self.total, n_terms, n_post, dev.index or 0))
        self._h = hnd
        N.check(N.lib().vqa_sparse_bind(hnd, ctypes.c_void_p(d["offsets"].data_ptr()),
                                        ctypes.c_void_p(d["docs"].data_ptr()), ctypes.c_void_p(weights.data_ptr())))
        self._ws = {}

    # ------------------------------------------------------------------ query
    def plan_queries(self, queries: Sequence[Any], limit: int):
        """Tokenise + classify each query's terms (Terms.search): returns the packed host arrays
        ``(q_terms int32[B,T], q_freqs float32[B,T], q_meta int32[B,4], k_cand_max)``."""
        lims = (ctypes.c_int32(), ctypes.c_int32())
        N.lib().vqa_sparse_limits(ctypes.byref(lims[0]), ctypes.byref(lims[1]))
        max_terms, max_cand = lims[0].value, lims[1].value
        n = self.total
        common_flag = self._common_flag()
        vocab_get = self.vocab.get
        t_rows, f_rows, meta, width, kmax = [], [], [], 1, 1
        for q in queries:
            rare_t, rare_f, com_t, com_f = [], [], [], []
            for term, freq in Counter(self.tokenize(q)).items():
                tid = vocab_get(term)
                if tid is None:
                    continue
                if common_flag[tid]:
                    com_t.append(tid)
                    com_f.append(freq)
                else:
                    rare_t.append(tid)
                    rare_f.append(freq)
            if not rare_t:  # only common terms: they are scored over all their documents
                rare_t, rare_f, com_t, com_f = com_t, com_f, [], []
            total = len(rare_t) + len(com_t)
            if total > max_terms:
                raise ValueError(f"a query may hold at most {max_terms} distinct indexed terms; got {total}")
            k_cand = max(1, min(n, limit * 5 if com_t else limit))
            if k_cand > max_cand:
                raise ValueError(f"limit {limit} needs {k_cand} sparse candidates; at most {max_cand} supported")
            t_rows.append(rare_t + com_t)
            f_rows.append(rare_f + com_f)
            meta.append((len(rare_t), len(com_t), k_cand, 0))
            width = max(width, total)
            kmax = max(kmax, k_cand)
        for tr, fr in zip(t_rows, f_rows):  # pad the rows: one array conversion instead of per-element stores
            pad = width - len(tr)
            if pad:
                tr.extend([-1] * pad)
                fr.extend([0] * pad)
        q_terms = np.asarray(t_rows, dtype=np.int32).reshape(len(t_rows), width)
        q_freqs = np.asarray(f_rows, dtype=np.float32).reshape(len(f_rows), width)
        q_meta = np.asarray(meta, dtype=np.int32).reshape(len(meta), _META_STRIDE)
        return q_terms, q_freqs, q_meta, kmax

    def _common_flag(self) -> List[bool]:
        """Per term: document frequency above ``cutoff * N`` (txtai ``Terms``: such terms are deferred)."""
        flags = getattr(self, "_common", None)
        if flags is None or len(flags) != len(self._df):
            flags = self._common = (self._df > self.cutoff * self.total).tolist()
        return flags

    def search_tensors(self, queries: Sequence[Any], limit: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-resident sparse search: ``(scores float64 [B, limit], positions int64 [B, limit])``;
        unused slots hold ``-inf`` / ``-1``."""
        if self._h is None or self.total == 0:
            raise RuntimeError("term index is empty: call index() or load() first")
        limit = int(limit)
        if limit < 1:
            raise ValueError(f"limit must be >= 1; got {limit}")
        queries = list(queries)
        if not queries:
            raise ValueError("no queries given")
        return self.search_planned(*self.plan_queries(queries, limit), limit)

    def search_planned(self, q_terms: np.ndarray, q_freqs: np.ndarray, q_meta: np.ndarray, k_cand_max: int,
                       limit: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """The device half of a search: one small H2D copy of the packed query plan, ``vqa_sparse_search``."""
        return self.launch_staged(self.stage_plan(q_terms, q_freqs, q_meta, k_cand_max, limit))

    def stage_plan(self, q_terms: np.ndarray, q_freqs: np.ndarray, q_meta: np.ndarray, k_cand_max: int, limit: int):
        """Upload a query plan (one H2D copy) and allocate workspace / outputs; returns the launch arguments."""
        k_cand_max = max(k_cand_max, min(limit, self.total))
        dev = self.device
        b, width = q_terms.shape
        lim = min(limit, k_cand_max)
        packed = np.concatenate([q_terms.view(np.uint8).ravel(), q_freqs.view(np.uint8).ravel(),
                                 q_meta.view(np.uint8).ravel()])
        blob = torch.from_numpy(packed).to(dev)
        key = (b, k_cand_max)
        ws = self._ws.get(key)
        if ws is None:
            need = ctypes.c_size_t()
            N.check(N.lib().vqa_sparse_workspace_bytes(self._h, b, k_cand_max, ctypes.byref(need)))
            ws = self._ws[key] = torch.empty(max(need.value, 8), dtype=torch.uint8, device=dev)
        out_s = torch.empty((b, lim), dtype=torch.float64, device=dev)
        out_i = torch.empty((b, lim), dtype=torch.int64, device=dev)
        return {"blob": blob, "off_f": q_terms.nbytes, "off_m": q_terms.nbytes + q_freqs.nbytes, "width": width,
                "b": b, "k_cand_max": k_cand_max, "lim": lim, "limit": limit, "ws": ws, "out_s": out_s,
                "out_i": out_i}

    def launch_staged(self, st) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = self.device
        base = st["blob"].data_ptr()
        out_s, out_i, b, lim, limit = st["out_s"], st["out_i"], st["b"], st["lim"], st["limit"]
        normalize = bool(self.normalize and self.avgscore)
        N.check(N.lib().vqa_sparse_search(
            self._h, ctypes.c_void_p(base), ctypes.c_void_p(base + st["off_f"]), ctypes.c_void_p(base + st["off_m"]),
            st["width"], b, st["k_cand_max"], lim, int(normalize), float(self.avgscore or 0.0),
            ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()), ctypes.c_void_p(st["ws"].data_ptr()),
            st["ws"].numel(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        if lim < limit:
            pad_s = torch.full((b, limit - lim), float("-inf"), dtype=torch.float64, device=dev)
            pad_i = torch.full((b, limit - lim), -1, dtype=torch.int64, device=dev)
            out_s, out_i = torch.cat([out_s, pad_s], 1), torch.cat([out_i, pad_i], 1)
        return out_s, out_i

    def batchsearch(self, queries: Sequence[Any], limit: int = 3) -> List[List[Tuple[int, float]]]:
        s, i = self.search_tensors(queries, limit)
        s_h, i_h = s.cpu().tolist(), i.cpu().tolist()
        return [[(p, sc) for p, sc in zip(ir, sr) if p >= 0] for sr, ir in zip(s_h, i_h)]

    def search(self, query: Any, limit: int = 3) -> List[Tuple[int, float]]:
        return self.batchsearch([query], limit)[0]

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        """``<path>.npz`` (CSR postings, frequencies, lengths) + ``<path>.terms.json`` (vocabulary, config)."""
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        np.savez(path + ".npz", **self._host)
        cfg = {k: v for k, v in self.config.items() if not callable(v)}
        with open(path + ".terms.json", "w", encoding="utf-8") as f:
            json.dump({"version": 1, "config": cfg, "total": self.total,
                       "vocab": sorted(self.vocab, key=self.vocab.get)}, f, ensure_ascii=False)

    def load(self, path: str) -> None:
        with open(path + ".terms.json", "r", encoding="utf-8") as f:
            meta = json.load(f)
        with np.load(path + ".npz", allow_pickle=False) as z:
            self._host = {k: z[k] for k in z.files}
        self.vocab = {t: i for i, t in enumerate(meta["vocab"])}
        self.total = int(meta["total"])
        self._stats()
        self._upload()

    def exists(self, path: str) -> bool:
        return os.path.exists(path + ".terms.json") and os.path.exists(path + ".npz")
