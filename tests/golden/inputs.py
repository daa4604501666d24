"""Seeded inputs that are too large to store; shared by make_golden.py and the tests."""
import numpy as np

D = 768


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def kat2():
    """10k x 768 docs with rows 17, 4711, 9999 identical; the query equals them."""
    docs = _unit(np.random.default_rng(2), 10000, D)
    docs[4711] = docs[17]
    docs[9999] = docs[17]
    return docs, docs[17:18].copy()


def config_a():
    """BASELINE.json configs[0]: 10k x 768 fp32 docs (seed 1234), 64 queries (seed 4321)."""
    return _unit(np.random.default_rng(1234), 10000, D), _unit(np.random.default_rng(4321), 64, D)
