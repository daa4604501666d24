"""Seeded synthetic corpora for the sparse-leg tests (token lists; Zipf-distributed vocabulary so
that some terms exceed the 10 % document-frequency cutoff and most do not)."""
import numpy as np


def zipf_corpus(n_docs, vocab, seed, min_len=4, max_len=40, a=1.3):
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    p = ranks ** (-a)
    p /= p.sum()
    lens = rng.integers(min_len, max_len + 1, size=n_docs)
    flat = rng.choice(vocab, size=int(lens.sum()), p=p)
    docs, at = [], 0
    for ln in lens:
        docs.append([f"w{t}" for t in flat[at:at + ln]])
        at += ln
    return docs


def queries_from(docs, n_queries, seed, vocab, n_terms=(1, 6)):
    """Each query mixes tokens drawn from a random document (so rare terms hit), the most common
    terms (over the cutoff) and an unknown token now and then; tokens may repeat."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_queries):
        d = docs[int(rng.integers(len(docs)))]
        k = int(rng.integers(n_terms[0], n_terms[1] + 1))
        q = [d[int(rng.integers(len(d)))] for _ in range(k)]
        mode = int(rng.integers(4))
        if mode == 0:
            q.append("w0")                      # the most common term
        elif mode == 1:
            q = ["w0", "w1"][:max(1, k % 3)]    # common terms only
        elif mode == 2:
            q.append("unknown-token")
            q.append(q[0])                      # repeated term -> freq 2
        out.append(q)
    return out


def corpus_small():
    docs = zipf_corpus(400, 300, seed=11)
    return docs, queries_from(docs, 24, seed=12, vocab=300)
