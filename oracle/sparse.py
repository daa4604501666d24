"""CPU oracle for the sparse (BM25) leg and the hybrid fusion -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this
module; the product package never does.

PARITY UNPINNED (same reason as ``oracle.py``): the reference builds its indexes with
``txtai.Embeddings(hybrid=True, ...)`` (inference_pipeline/db_utils/heavy_ranker.py:78-83) and
everything behind that flag -- tokeniser, BM25 statistics, the term index, score
normalisation and the dense/sparse fusion -- lives in third-party ``txtai``
(``requirements.txt:74``, unpinned, un-vendored, not installable here).  This file restates
the published behaviour of the txtai 6.x line class by class:

* ``tokenize``        -- txtai ``Tokenizer()`` as the scoring index uses it when ``terms=True``:
                         lower-case, Unicode (UAX #29) word segmentation, no stop words.
* ``BM25``            -- txtai ``scoring.TFIDF`` / ``scoring.BM25``: document / word
                         frequencies, ``idf = log(1 + (N - n + 0.5) / (n + 0.5))``,
                         ``score = idf * f * (k1 + 1) / (f + k1 * (1 - b + b * dl / avgdl))``
                         (k1 = 1.2, b = 0.75), ``avgscore`` and the ``normalize=True`` rule
                         ``min(score / min(top + avgscore, 6 * avgscore), 1.0)``.
* ``BM25.search``     -- txtai ``scoring.Terms.search``: a dense fp32 accumulator, terms taken in
                         the query's first-occurrence order, terms present in more than
                         ``cutoff`` (10 %) of the documents deferred and merged only into the
                         ``5 * limit`` best candidates of the rarer terms, results with score 0
                         dropped.
* ``hybrid``          -- txtai ``Search``: both legs fetch ``10 * limit`` candidates, scores are
                         added per id with weights ``[w, 1 - w]`` (w = 0.5) in Python floats, and a
                         stable descending sort keeps the first ``limit``.

One deliberate definition: txtai selects candidates with ``np.argpartition`` and orders them
with an unstable ``argsort``, so which of several EQUAL scores it returns is implementation
defined.  As everywhere else in this repository the order is pinned to *score descending, ties
-> lower document position* (BASELINE.json north_star).
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_SEGMENT = None


def tokenize(text: str) -> List[str]:
    """txtai ``Tokenizer(lowercase=True, emoji=True, alphanum=False, stopwords=False)``."""
    global _SEGMENT
    if _SEGMENT is None:
        import regex

        _SEGMENT = regex.compile(r"[\w\p{Extended_Pictographic}\p{WB:RegionalIndicator}](?:\B\S)*", flags=regex.WORD)
    return _SEGMENT.findall(text.lower())


class BM25:
    """Sparse keyword index over document positions 0..N-1."""

    def __init__(self, k1: float = 1.2, b: float = 0.75, cutoff: float = 0.1, normalize: bool = True):
        self.k1, self.b, self.cutoff, self.normalize = k1, b, cutoff, normalize
        self.total = 0
        self.docfreq: Counter = Counter()
        self.wordfreq: Counter = Counter()
        self.lengths: List[int] = []
        self.postings: Dict[str, Tuple[List[int], List[int]]] = {}
        self.idf: Dict[str, float] = {}
        self.tokens = 0
        self.avgfreq = self.avgdl = self.avgidf = 0.0
        self.avgscore: Optional[float] = None

    # -- build (TFIDF.insert / addstats / index, Terms.insert) -------------------------------
    def index(self, documents: Iterable[Sequence[str]]) -> "BM25":
        """``documents``: one token list per document, in position order."""
        for uid, tokens in enumerate(documents):
            self.lengths.append(len(tokens))
            for term, freq in Counter(tokens).items():
                uids, freqs = self.postings.setdefault(term, ([], []))
                uids.append(uid)
                freqs.append(freq)
            self.wordfreq.update(tokens)
            self.docfreq.update(list(dict.fromkeys(tokens)))  # txtai: set(tokens); order pinned to first occurrence
            self.total += 1
        if self.wordfreq:
            self.tokens = sum(self.wordfreq.values())
            self.avgfreq = self.tokens / len(self.wordfreq.values())
            self.avgdl = self.tokens / self.total
            idfs = self.computeidf(np.array(list(self.docfreq.values())))
            for x, word in enumerate(self.docfreq):
                self.idf[word] = float(idfs[x])
            self.avgidf = float(np.mean(idfs))
            self.avgscore = float(self.score(self.avgfreq, self.avgidf, self.avgdl))
        self._lengths = np.array(self.lengths, dtype=np.int64)
        return self

    def computeidf(self, freq):
        return np.log(1 + (self.total - freq + 0.5) / (freq + 0.5))

    def score(self, freq, idf, length):
        k = self.k1 * ((1 - self.b) + self.b * length / self.avgdl)
        return idf * (freq * (self.k1 + 1)) / (freq + k)

    def weights(self, term: str):
        """Terms.weights: (positions int64[], fp32 weight per posting) or (None, None)."""
        if term not in self.postings:
            return None, None
        uids = np.array(self.postings[term][0], dtype=np.int64)
        freqs = np.array(self.postings[term][1], dtype=np.int64)
        w = self.score(freqs, self.idf[term], self._lengths[uids]).astype(np.float32)
        return uids, w

    # -- query (Terms.search / topn / merge, TFIDF.search) -----------------------------------
    @staticmethod
    def _best(scores: np.ndarray, n: int) -> np.ndarray:
        """The n best positions: score descending, ties -> lower position."""
        return np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))[:n]

    def raw_search(self, terms: Sequence[str], limit: int) -> List[Tuple[int, float]]:
        n = self.total
        scores = np.zeros(n, dtype=np.float32)
        counted, skipped, hasscores = Counter(terms), {}, False
        for term, freq in counted.items():
            uids, w = self.weights(term)
            if uids is not None:
                if len(uids) <= self.cutoff * n:
                    scores[uids] += np.float32(freq) * w
                    hasscores = True
                else:
                    skipped[term] = freq
        topn = min(n, limit * 5 if skipped else limit)
        matches = self._best(scores, topn)
        for term, freq in skipped.items():
            uids, w = self.weights(term)
            if hasscores:
                idx = np.searchsorted(uids, matches)
                idx = np.array([x for i, x in enumerate(idx) if x < len(uids) and uids[x] == matches[i]], dtype=np.int64)
                uids, w = uids[idx], w[idx]
            scores[uids] += np.float32(freq) * w
        if not hasscores:
            matches = self._best(scores, topn)
        order = np.lexsort((matches, -scores[matches].astype(np.float64)))
        matches = matches[order]
        return [(int(x), float(scores[x])) for x in matches[:limit] if scores[x] > 0]

    def search(self, query, limit: int = 3) -> List[Tuple[int, float]]:
        terms = tokenize(query) if isinstance(query, str) else list(query)
        scores = self.raw_search(terms, limit)
        if self.normalize and scores:
            maxscore = min(scores[0][1] + self.avgscore, 6 * self.avgscore)
            scores = [(x, min(score / maxscore, 1.0)) for x, score in scores]
        return scores


def hybrid(dense: List[Tuple[int, float]], sparse: List[Tuple[int, float]], limit: int, weights=0.5,
           normalized: bool = True):
    """txtai ``Search`` fusion of one query's dense and sparse candidate lists (recalled txtai 6.x
    embeddings/search/base.py): weighted score sum when the scoring index is normalised,
    reciprocal-rank fusion ``1 / (rank + 1) * weight`` otherwise; a leg whose weight is not > 0 is skipped."""
    if isinstance(weights, (int, float)):
        weights = [weights, 1 - weights]
    uids: Dict[int, float] = {}
    for v, scores in enumerate((dense, sparse)):
        for r, (uid, score) in enumerate(scores if weights[v] > 0 else []):
            if uid not in uids:
                uids[uid] = 0.0
            if normalized:
                uids[uid] += score * weights[v]
            else:
                uids[uid] += (1.0 / (r + 1)) * weights[v]
    return sorted(uids.items(), key=lambda x: x[1], reverse=True)[:limit]
