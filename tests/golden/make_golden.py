"""Generates tests/golden/kat.npz -- the known-answer vectors for the retrieval hot path.

The reference holds no golden vectors (parity unpinned) and its arithmetic
(txtai -> faiss) cannot be imported in this container, so these are AUTHORED here: inputs
are constructed so that the right answer is known by construction, and the stored outputs
come from the numpy restatement of the txtai NumPy backend (oracle.np_search: np.dot +
stable descending sort), which is independent of the C oracle and of the CUDA path.

    python tests/golden/make_golden.py        # rewrites kat.npz (deterministic)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tests.golden import inputs  # noqa: E402

out = {}
D = 768

# KAT-1 identity: docs = I_8 padded to D=768, q = e_3 -> [3, then 0,1,2,4..] by the tie rule
docs = np.zeros((8, D), np.float32)
docs[np.arange(8), np.arange(8)] = 1.0
q = np.zeros((1, D), np.float32)
q[0, 3] = 1.0
out["kat1_docs"], out["kat1_q"] = docs, q
out["kat1_scores"], out["kat1_ids"] = oracle.np_search(docs, q, 8)
assert out["kat1_ids"][0].tolist() == [3, 0, 1, 2, 4, 5, 6, 7]

# KAT-2 planted duplicates: rows 17, 4711, 9999 identical and equal to the query -> ascending id
docs, q = inputs.kat2()  # regenerated from a seed by the tests (30 MB, not stored)
out["kat2_scores"], out["kat2_ids"] = oracle.np_search(docs, q, 5)
assert out["kat2_ids"][0, :3].tolist() == [17, 4711, 9999]

# KAT-3 exact scores {-1,-0.5,0,0.5,1}, exactly representable in fp32/bf16/fp16
docs = np.zeros((5, D), np.float32)
docs[0, 0] = -1.0
docs[1, 0], docs[1, 1] = -0.5, 0.5
docs[2, 1] = 1.0
docs[3, 0], docs[3, 1] = 0.5, -0.5
docs[4, 0] = 1.0
q = np.zeros((1, D), np.float32)
q[0, 0] = 1.0
out["kat3_docs"], out["kat3_q"] = docs, q
out["kat3_scores"], out["kat3_ids"] = oracle.np_search(docs, q, 5)
assert out["kat3_scores"][0].tolist() == [1.0, 0.5, 0.0, -0.5, -1.0]
assert out["kat3_ids"][0].tolist() == [4, 3, 2, 1, 0]

# KAT-4 k == N and k > N (padding with (-inf, -1))
rng = np.random.default_rng(4)
docs = rng.standard_normal((37, 384)).astype(np.float32)
docs /= np.linalg.norm(docs, axis=1, keepdims=True)
q = rng.standard_normal((3, 384)).astype(np.float32)
q /= np.linalg.norm(q, axis=1, keepdims=True)
out["kat4_docs"], out["kat4_q"] = docs, q
out["kat4_scores_kN"], out["kat4_ids_kN"] = oracle.np_search(docs, q, 37)
out["kat4_scores_k64"], out["kat4_ids_k64"] = oracle.np_search(docs, q, 64)

# KAT-5 duplicates straddling shard boundaries (2-, 4-, 8-way row shards of 1024 rows)
rng = np.random.default_rng(5)
docs = rng.standard_normal((1024, 384)).astype(np.float32)
docs /= np.linalg.norm(docs, axis=1, keepdims=True)
for r in (127, 128, 511, 512, 767, 768, 1023):
    docs[r] = docs[0]
q = docs[0:1].copy()
out["kat5_docs"], out["kat5_q"] = docs, q
out["kat5_scores"], out["kat5_ids"] = oracle.np_search(docs, q, 10)
assert out["kat5_ids"][0, :8].tolist() == [0, 127, 128, 511, 512, 767, 768, 1023]

# KAT-6 mean-pool: all-padding row (denominator clamp), S = 1, ragged lengths
rng = np.random.default_rng(6)
h = rng.standard_normal((4, 9, 64)).astype(np.float32)
m = np.zeros((4, 9), np.int64)
m[0, :9] = 1
m[1, :1] = 1
m[2, :] = 0
m[3, :5] = 1
out["kat6_hidden"], out["kat6_mask"] = h, m
out["kat6_pooled"] = oracle.np_mean_pool(h, m, normalize=False)
out["kat6_pooled_norm"] = oracle.np_mean_pool(h, m, normalize=True)
assert np.all(out["kat6_pooled"][2] == 0) and np.all(out["kat6_pooled_norm"][2] == 0)

# KAT-7 zero-norm row in normalise
x = rng.standard_normal((6, 64)).astype(np.float32)
x[3] = 0
out["kat7_x"] = x
out["kat7_norm"] = oracle.np_normalize_rows(x)

# Config A of BASELINE.json at golden size: 10k x 768 fp32, 64 queries, top-5 (seeded)
docsA, qA = inputs.config_a()
sA, iA = oracle.np_search(docsA, qA, 5)
out["cfgA_ids"], out["cfgA_scores"] = iA, sA  # inputs are regenerated from the seeds by the tests

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.npz"), **out)
print("wrote kat.npz:", {k: v.shape for k, v in out.items()})
