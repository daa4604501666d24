"""One timing run; knobs come from the environment (ROWS, BATCHES, MODE, K, VQA_*)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200 import ops  # noqa: E402

n, d = int(os.environ.get("ROWS", "4000000")), int(os.environ.get("DIM", "768"))
k = int(os.environ.get("K", "10"))
dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[os.environ.get("DTYPE", "bf16")]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
rows = torch.empty((n, d), dtype=dt, device=dev)
for lo in range(0, n, 500000):
    m = min(500000, n - lo)
    x = ops.normalize_rows(torch.randn((m, d), generator=g, device=dev))
    rows[lo:lo + m] = x.to(dt)
shard = ops.FlatShard(rows)
res = {}
iters = int(os.environ.get("ITERS", "20"))
for b in [int(x) for x in os.environ.get("BATCHES", "8,32").split(",")]:
    q = ops.normalize_rows(torch.randn((b, d), generator=g, device=dev))
    mode = os.environ.get("MODE", "tensor")
    for _ in range(3):
        shard.search(q, k, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        shard.search(q, k, mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res[b] = {"ms": round(ms, 4), "GBps": round(n * d * rows.element_size() / ms / 1e6)}
    if os.environ.get("CHECK"):  # recall@k and score error against the fp32 verify kernel on the same rows
        s1, i1 = shard.search(q, k, mode)
        s0, i0 = shard.search(q, k, "verify")
        torch.cuda.synchronize()
        a, r = i1.cpu().tolist(), i0.cpu().tolist()
        res[b]["recall"] = round(sum(len(set(x) & set(y)) for x, y in zip(a, r)) / (len(r) * k), 5)
        res[b]["max_rel_err"] = float(((s1 - s0).abs() / s0.abs().clamp_min(1e-3)).max())
print(json.dumps(res))
