"""The CPU oracle against the authored golden vectors and against its own two restatements
(C canonical/semantic vs numpy).  CPU only."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from tests.conftest import unit_rows
from tests.golden import inputs


def both(docs, q, k, **kw):
    return oracle.search(docs, q, k, oracle.CANONICAL, **kw), oracle.search(docs, q, k, oracle.SEMANTIC, **kw)


def test_kat1_identity(golden):
    (s, i), (s2, i2) = both(golden["kat1_docs"], golden["kat1_q"], 8)
    assert i.tolist() == golden["kat1_ids"].tolist() == i2.tolist() == [[3, 0, 1, 2, 4, 5, 6, 7]]
    assert s.tolist() == golden["kat1_scores"].tolist() == [[1, 0, 0, 0, 0, 0, 0, 0]]


def test_kat2_planted_duplicates_ascending_id(golden):
    docs, q = inputs.kat2()
    (s, i), (_, i2) = both(docs, q, 5)
    assert i[0, :3].tolist() == [17, 4711, 9999]
    assert np.array_equal(i, golden["kat2_ids"]) and np.array_equal(i2, golden["kat2_ids"])
    assert s[0, 0] == s[0, 1] == s[0, 2]
    np.testing.assert_allclose(s, golden["kat2_scores"], rtol=1e-5)


def test_kat3_exact_scores(golden):
    for storage in ("fp32", "bf16", "fp16"):
        s, i = oracle.search(golden["kat3_docs"], golden["kat3_q"], 5, oracle.CANONICAL, storage)
        assert s.tolist() == [[1.0, 0.5, 0.0, -0.5, -1.0]] and i.tolist() == [[4, 3, 2, 1, 0]]


def test_kat4_k_equals_n_and_k_exceeds_n(golden):
    (s, i), _ = both(golden["kat4_docs"], golden["kat4_q"], 37)
    assert np.array_equal(i, golden["kat4_ids_kN"])
    (s, i), _ = both(golden["kat4_docs"], golden["kat4_q"], 64)
    assert np.array_equal(i, golden["kat4_ids_k64"])
    assert np.all(i[:, 37:] == -1) and np.all(np.isneginf(s[:, 37:]))


def test_kat5_shard_boundary_ties_and_merge(golden):
    docs, q = golden["kat5_docs"], golden["kat5_q"]
    s, i = oracle.search(docs, q, 10)
    assert np.array_equal(i, golden["kat5_ids"])
    for g in (2, 4, 8):
        per = 1024 // g
        parts = [oracle.search(docs[r * per:(r + 1) * per], q, 10, first_id=r * per) for r in range(g)]
        ms, mi = oracle.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 10)
        assert np.array_equal(mi, i) and np.array_equal(ms, s)


def test_kat6_mean_pool(golden):
    h, m = golden["kat6_hidden"], golden["kat6_mask"]
    np.testing.assert_allclose(oracle.mean_pool(h, m, False), golden["kat6_pooled"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(oracle.mean_pool(h, m, True), golden["kat6_pooled_norm"], rtol=1e-5, atol=1e-6)
    assert np.all(oracle.mean_pool(h, m, True)[2] == 0)            # all-padding row -> clamp -> zeros
    np.testing.assert_allclose(oracle.mean_pool(h, m, False)[1], h[1, 0], rtol=1e-6)  # S_valid = 1


def test_kat7_zero_norm_row(golden):
    out = oracle.normalize_rows(golden["kat7_x"])
    np.testing.assert_allclose(out, golden["kat7_norm"], rtol=1e-6, atol=1e-7)
    assert np.all(out[3] == 0) and np.all(np.isfinite(out))


def test_config_a_matches_numpy_backend(golden):
    docs, q = inputs.config_a()
    (s, i), (s2, i2) = both(docs, q, 5)
    assert np.array_equal(i, golden["cfgA_ids"]) and np.array_equal(i2, golden["cfgA_ids"])
    np.testing.assert_allclose(s, golden["cfgA_scores"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(s2, golden["cfgA_scores"], rtol=1e-5, atol=1e-7)
    fs, fi = oracle.np_search_fast(docs, q, 5)
    assert np.array_equal(fi, golden["cfgA_ids"])


def test_canonical_within_1e5_of_semantic():
    rng = np.random.default_rng(0)
    docs, q = unit_rows(rng, 2000, 768), unit_rows(rng, 8, 768)
    a = oracle.scores(docs, q, oracle.CANONICAL)
    b = oracle.scores(docs, q, oracle.SEMANTIC)
    assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()


def test_canonical_order_is_what_the_header_says():
    """Replay SURVEY.md App. C by hand (python floats -> fp32 via numpy) for one row."""
    rng = np.random.default_rng(3)
    for d, e in ((768, 4), (768, 8), (200, 4), (72, 8)):
        q, x = rng.standard_normal(d).astype(np.float32), rng.standard_normal(d).astype(np.float32)
        part = np.zeros(32, np.float32)
        nch = (d + e - 1) // e
        for lane in range(32):
            acc = np.float32(0)
            for c in range(lane, nch, 32):
                for j in range(c * e, min(c * e + e, d)):
                    acc = np.float32(np.float64(q[j]) * np.float64(x[j]) + np.float64(acc))  # fma: one rounding
            part[lane] = acc
        for m in (16, 8, 4, 2, 1):
            part = (part + part[np.arange(32) ^ m]).astype(np.float32)
        got = oracle.lib().oracle_dot_canonical(q.ctypes.data_as(oracle.oracle.ctypes.POINTER(oracle.oracle.ctypes.c_float)),
                                                x.ctypes.data_as(oracle.oracle.ctypes.POINTER(oracle.oracle.ctypes.c_float)), d, e)
        assert np.float32(got) == part[0]


def test_agreement_rule():
    assert oracle.agree(5, 0.25, 5, 0.2) is True          # 0.45 > 0.4
    # scores are fp32 values turned into Python floats: fp32(0.2)+fp32(0.2) = 0.4000000059... > 0.4
    assert oracle.agree(5, 0.2, 5, 0.2) is True
    assert oracle.agree(5, 0.125, 5, 0.25) is False       # 0.375 exactly
    assert oracle.agree(5, 0.3, 6, 0.3) is False          # different ids
    assert oracle.agree(7, 0.1, 7, 0.1) is False


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 300), st.integers(1, 12), st.integers(0, 2 ** 31 - 1))
def test_property_permutation_and_shard_merge(n, k, seed):
    rng = np.random.default_rng(seed)
    d = 64
    docs, q = unit_rows(rng, n, d), unit_rows(rng, 2, d)
    s, i = oracle.search(docs, q, k)
    # (1) agrees with the numpy backend restatement
    s2, i2 = oracle.np_search(docs, q, k)
    fin = i >= 0
    assert np.array_equal(fin, i2 >= 0)
    np.testing.assert_allclose(s[fin], s2[fin], rtol=1e-5, atol=1e-6)
    # (2) concatenating shards == merge of shard results (any split point)
    cut = int(rng.integers(0, n + 1))
    a = oracle.search(docs[:cut], q, k) if cut > 0 else (np.full((2, k), -np.inf, np.float32), np.full((2, k), -1, np.int64))
    b = oracle.search(docs[cut:], q, k, first_id=cut) if cut < n else (np.full((2, k), -np.inf, np.float32), np.full((2, k), -1, np.int64))
    ms, mi = oracle.merge_topk(np.stack([a[0], b[0]]), np.stack([a[1], b[1]]), k)
    assert np.array_equal(mi, i) and np.array_equal(ms, s)
    # (3) permuting rows permutes ids consistently (scores are order independent per row)
    perm = rng.permutation(n)
    sp, ip = oracle.search(docs[perm], q, k)
    assert np.array_equal(np.sort(sp, axis=1), np.sort(s, axis=1))


@pytest.mark.parametrize("n,k", [(3000, 129), (5000, 300), (2500, 1000), (900, 1000)])
def test_wide_k_agrees_with_the_numpy_restatement(n, k):
    """k > 128 (what a hybrid search with limit > 12 asks of the dense leg): the C oracle that pins the GPU's composed
    wide top-k (tests/test_gpu_search.py::test_wide_k_*) against the numpy-backend restatement -- same ids wherever
    the scores are not within rounding of each other, planted exact duplicates lower id first, padding beyond n."""
    rng = np.random.default_rng(n + k)
    d = 96
    docs, q = unit_rows(rng, n, d), unit_rows(rng, 3, d)
    docs[n // 2] = docs[7]
    docs[n - 1] = docs[7]
    q[0] = docs[7]
    s, i = oracle.search(docs, q, k)
    s2, i2 = oracle.np_search(docs, q, k)
    kk = min(k, n)
    assert np.all(i[:, kk:] == -1) and np.all(np.isneginf(s[:, kk:])) and np.array_equal(i >= 0, i2 >= 0)
    assert i[0, :3].tolist() == [7, n // 2, n - 1]
    np.testing.assert_allclose(s[:, :kk], s2[:, :kk], rtol=1e-5, atol=1e-6)
    assert np.all(np.diff(s[:, :kk], axis=1) <= 0)
    differ = i[:, :kk] != i2[:, :kk]                  # only where neighbouring scores are within fp32 rounding
    for r, c in zip(*np.nonzero(differ)):
        assert abs(float(s2[r, c]) - float(s[r, c])) <= 2e-6
    assert differ.mean() < 0.01
