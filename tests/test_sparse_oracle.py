"""The sparse-leg oracle (oracle/sparse.py) against hand-computed known answers, and the product's
HOST logic (tokeniser, CSR postings, query planning) against the oracle.  CPU only: nothing here
scores on the device."""
import math
from collections import Counter

import numpy as np
import pytest

from oracle import sparse as osp
from tests.golden import sparse_inputs as si
from vietnamese_qa_system_b200.scoring import BM25, Tokenizer


# ---- tokeniser ---------------------------------------------------------------------------
TOKEN_KAT = [
    ("Xin chào, thế giới!", ["xin", "chào", "thế", "giới"]),
    ("HÀ NỘI là thủ đô của Việt Nam.", ["hà", "nội", "là", "thủ", "đô", "của", "việt", "nam"]),
    ("Don't stop 3.14", ["don't", "stop", "3.14"]),
    ("", []),
    ("  ...  ", []),
]


@pytest.mark.parametrize("text,want", TOKEN_KAT)
def test_tokenizer_known_answers(text, want):
    assert osp.tokenize(text) == want
    assert Tokenizer()(text) == want          # product tokeniser == oracle tokeniser


def test_tokenizer_options():
    assert Tokenizer(alphanum=True, stopwords=True)("The Cat-5 cable, and a 9 to 5 job") == ["cat-5", "cable", "job"]
    assert Tokenizer(lowercase=False)("Hà Nội") == ["Hà", "Nội"]
    assert Tokenizer(stopwords=["nội"])("Hà Nội") == ["hà"]


# ---- BM25 known answers ----------------------------------------------------------------------
def test_bm25_single_term_equal_lengths_scores_are_idf():
    # every document has avgdl tokens and tf = 1  =>  k = k1 and score = idf * (k1+1)/(1+k1) = idf
    docs = [["a", "b"], ["a", "c"], ["d", "e"], ["f", "g"]]
    bm = osp.BM25(normalize=False).index(docs)
    idf_a = math.log(1 + (4 - 2 + 0.5) / (2 + 0.5))
    idf_d = math.log(1 + (4 - 1 + 0.5) / (1 + 0.5))
    assert bm.search(["a"], 3) == [(0, float(np.float32(idf_a))), (1, float(np.float32(idf_a)))]   # tie -> lower id
    assert bm.search(["d"], 3) == [(2, float(np.float32(idf_d)))]
    assert bm.search(["zzz"], 3) == []                                                          # unknown term
    # repeated query term: weight * 2
    assert bm.search(["d", "d"], 1) == [(2, float(np.float32(2) * np.float32(idf_d)))]


def test_bm25_length_normalisation_by_hand():
    docs = [["x"], ["x", "y", "y", "y"], ["z"], ["z"], ["z"], ["z"], ["z"], ["z"], ["z"], ["z"], ["z"], ["z"],
            ["z"], ["z"], ["z"], ["z"], ["z"], ["z"], ["z"], ["z"]]
    bm = osp.BM25(normalize=False).index(docs)
    n, avgdl = 20, (1 + 4 + 18) / 20
    idf = math.log(1 + (n - 2 + 0.5) / (2 + 0.5))

    def w(f, dl):
        k = 1.2 * ((1 - 0.75) + 0.75 * dl / avgdl)
        return float(np.float32(idf * (f * (1.2 + 1)) / (f + k)))

    assert bm.search(["x"], 5) == [(0, w(1, 1)), (1, w(1, 4))]     # the shorter document scores higher
    assert w(1, 1) > w(1, 4)


def test_common_terms_are_deferred_and_merged_into_candidates_only():
    docs, queries = si.corpus_small()
    bm = osp.BM25(normalize=False).index(docs)
    n = len(docs)
    common = [t for t, uids in bm.postings.items() if len(uids[0]) > 0.1 * n]
    rare = [t for t, uids in bm.postings.items() if len(uids[0]) <= 0.1 * n]
    assert common and rare
    # a common term next to a rare one only re-scores the rare term's candidates (+ zero-score fill)
    res = bm.raw_search([rare[0], common[0]], 2)
    rare_docs = set(bm.postings[rare[0]][0])
    assert res and res[0][0] in rare_docs
    # only common terms: scored over all their documents
    res2 = bm.raw_search([common[0]], 3)
    assert len(res2) == 3 and all(p in set(bm.postings[common[0]][0]) for p, _ in res2)


def test_normalize_rule():
    docs, _ = si.corpus_small()
    raw = osp.BM25(normalize=False).index(docs)
    nrm = osp.BM25(normalize=True).index(docs)
    q = docs[3][:3]
    r, m = raw.search(q, 5), nrm.search(q, 5)
    maxscore = min(r[0][1] + raw.avgscore, 6 * raw.avgscore)
    assert [p for p, _ in r] == [p for p, _ in m]
    assert [s for _, s in m] == [min(s / maxscore, 1.0) for _, s in r]
    assert all(0 < s <= 1.0 for _, s in m)


def test_hybrid_fusion_by_hand():
    dense = [(5, 0.9), (7, 0.5), (2, 0.5)]
    sparse = [(7, 0.8), (9, 0.5), (5, 0.0625)]
    got = osp.hybrid(dense, sparse, 4)
    assert got == [(7, 0.5 * 0.5 + 0.8 * 0.5), (5, 0.9 * 0.5 + 0.0625 * 0.5), (2, 0.25), (9, 0.25)]  # 2 before 9: insertion order
    assert osp.hybrid(dense, sparse, 2, weights=1.0) == [(5, 0.9), (7, 0.5)]
    assert osp.hybrid(dense, [], 2) == [(5, 0.45), (7, 0.25)]


# ---- product host logic (no device) --------------------------------------------------------
def test_csr_postings_and_statistics_match_oracle():
    docs, _ = si.corpus_small()
    bm = BM25({"terms": True, "normalize": True})
    bm.build_postings(docs)
    ref = osp.BM25().index(docs)
    h = bm._host
    assert bm.total == ref.total and bm.tokens == ref.tokens
    assert bm.avgdl == ref.avgdl and bm.avgfreq == ref.avgfreq
    assert bm.avgidf == ref.avgidf and bm.avgscore == ref.avgscore          # bit-identical scalars
    assert list(bm.vocab) == list(ref.docfreq)                             # same first-occurrence order
    for term, tid in bm.vocab.items():
        lo, hi = h["offsets"][tid], h["offsets"][tid + 1]
        assert h["docs"][lo:hi].tolist() == ref.postings[term][0]
        assert h["freqs"][lo:hi].tolist() == ref.postings[term][1]
        assert bm.idf_host[tid] == ref.idf[term]
    assert h["lengths"].tolist() == ref.lengths
    assert np.all(np.diff(h["offsets"]) > 0)


def test_query_planning_matches_terms_search_classification():
    docs, queries = si.corpus_small()
    bm = BM25({"terms": True})
    bm.build_postings(docs)
    ref = osp.BM25().index(docs)
    n = len(docs)
    q_terms, q_freqs, q_meta, kmax = bm.plan_queries(queries, 4)
    inv = {v: k for k, v in bm.vocab.items()}
    for i, q in enumerate(queries):
        counted = Counter(q)
        known = [(t, f) for t, f in counted.items() if t in ref.postings]
        rare = [(t, f) for t, f in known if len(ref.postings[t][0]) <= 0.1 * n]
        common = [(t, f) for t, f in known if len(ref.postings[t][0]) > 0.1 * n]
        if not rare:
            rare, common = common, []
        n_rare, n_common, k_cand = q_meta[i, :3]
        assert (n_rare, n_common) == (len(rare), len(common))
        assert k_cand == min(n, 20 if common else 4)
        got = [(inv[t], f) for t, f in zip(q_terms[i, :n_rare + n_common].tolist(),
                                           q_freqs[i, :n_rare + n_common].tolist())]
        assert got == [(t, float(f)) for t, f in rare + common]
        assert np.all(q_terms[i, n_rare + n_common:] == -1)
    assert kmax == int(q_meta[:, 2].max())


def test_query_planning_limits():
    bm = BM25({"terms": True})
    bm.build_postings([[f"t{i}" for i in range(100)]] * 30)     # every term is common (df = N)
    with pytest.raises(ValueError):
        bm.plan_queries([[f"t{i}" for i in range(65)]], 3)       # > 64 distinct terms
    bm.plan_queries([[f"t{i}" for i in range(64)]], 3)
    bm2 = BM25({"terms": True})
    bm2.build_postings([["r%d" % i, "c"] for i in range(3000)])
    with pytest.raises(ValueError):
        bm2.plan_queries([["r1", "c"]], 300)                     # 5 x 300 candidates > 1024


# ---- the oracle against the textbook definition (independent code path) ---------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


def definitional_scores(docs, query, k1=1.2, b=0.75):
    """BM25 straight from the formula, document by document (no postings, no arrays): fp32 weight per
    (term, document), fp32 accumulation in the query's first-occurrence order."""
    n = len(docs)
    avgdl = sum(len(d) for d in docs) / n
    out = []
    for d in docs:
        s = np.float32(0.0)
        for term, qf in Counter(query).items():
            df = sum(1 for x in docs if term in x)
            tf = d.count(term)
            if df == 0 or tf == 0:
                continue
            idf = float(np.log(1 + (n - df + 0.5) / (df + 0.5)))
            k = k1 * ((1 - b) + b * len(d) / avgdl)
            w = np.float32(idf * (tf * (k1 + 1)) / (tf + k))
            s = np.float32(s + np.float32(np.float32(qf) * w))
        out.append(float(s))
    return out


@settings(max_examples=60, deadline=None, derandomize=True)
@given(st.integers(0, 10_000), st.integers(12, 60), st.integers(3, 12))
def test_oracle_equals_definition_when_no_term_is_deferred(seed, n_docs, vocab):
    rng = np.random.default_rng(seed)
    docs = [[f"t{int(x)}" for x in rng.integers(0, vocab, int(rng.integers(1, 9)))] for _ in range(n_docs)]
    query = [f"t{int(x)}" for x in rng.integers(0, vocab + 2, int(rng.integers(1, 5)))]
    bm = osp.BM25(normalize=False, cutoff=1.0).index(docs)          # cutoff 1.0: nothing is "common"
    want = definitional_scores(docs, query)
    order = sorted((i for i in range(n_docs) if want[i] > 0), key=lambda i: (-want[i], i))
    got = bm.search(query, n_docs)
    assert got == [(i, want[i]) for i in order]
    assert bm.search(query, 3) == [(i, want[i]) for i in order[:3]]


@settings(max_examples=40, deadline=None, derandomize=True)
@given(st.integers(0, 10_000))
def test_only_common_terms_are_scored_over_all_documents(seed):
    rng = np.random.default_rng(seed)
    docs = [["c"] * int(rng.integers(1, 4)) + [f"r{int(x)}" for x in rng.integers(0, 50, 3)] for _ in range(40)]
    bm = osp.BM25(normalize=False).index(docs)                       # "c" is in every document -> deferred
    want = definitional_scores(docs, ["c"])
    order = sorted(range(40), key=lambda i: (-want[i], i))
    assert bm.search(["c"], 7) == [(i, want[i]) for i in order[:7]]
