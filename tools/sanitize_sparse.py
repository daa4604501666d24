"""Small-shape run of the sparse (BM25) leg and the fusion kernels for compute-sanitizer.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_sparse.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200 import ops  # noqa: E402
from vietnamese_qa_system_b200.scoring import BM25  # noqa: E402

n = 40_000                                   # 3 score tiles
docs = []
for i in range(n):
    d = ["pad%d" % (i % 5)]
    if i % 2 == 0:
        d += ["all"] * (1 + i % 3)           # 50 %: common; alone in a query it is accumulated over dense tiles
    if i % 9 == 0:
        d.append("mid")                      # 11 %: common, deferred next to a rare term
    if i % 1000 == 7:
        d.append("rare")
    docs.append(d)
bm = BM25({"method": "bm25", "terms": True, "normalize": True})
bm.index(docs)
queries = [["all"], ["rare"], ["rare", "mid"], ["rare", "all", "mid"], ["all", "mid"], ["nothing"], ["pad1", "rare"]]
for limit in (1, 10, 60):
    s, i = bm.search_tensors(queries, limit)
    torch.cuda.synchronize()
    assert int((i[0] >= 0).sum()) == limit and int((i[5] >= 0).sum()) == 0
ds = torch.rand(7, 10, device="cuda").sort(dim=1, descending=True).values
di = torch.stack([torch.randperm(50, device="cuda")[:10] for _ in range(7)])
s, i = bm.search_tensors(queries, 10)
ops.hybrid_fuse(ds, di, s, i, 5)
ops.agree(i[:, 0].contiguous(), s[:, 0].contiguous(), i[:, 0].contiguous(), s[:, 0].contiguous(), 0.4)
torch.cuda.synchronize()
print("sanitize (sparse) run complete")
