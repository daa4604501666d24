"""One timing run; shape from the environment (ROWS, DIM, DTYPE, BATCHES, MODE, K), kernel variants as
VARIANTS="-;VQA_TS_KS=4;VQA_REDUCE_EARLY=0,VQA_TS_QS=0": each is applied to the ONE generated index through
FlatShard.set_tuning (the library reads its environment only in vqa_index_create; "-" = the handle as created)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200 import _native as N, ops  # noqa: E402

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
iters = int(os.environ.get("ITERS", "20"))
mode = os.environ.get("MODE", "tensor")
batches = [int(x) for x in os.environ.get("BATCHES", "8,32").split(",")]
queries = {b: ops.normalize_rows(torch.randn((b, d), generator=g, device=dev)) for b in batches}
variants = [v for v in os.environ.get("VARIANTS", "-").split(";") if v]
base = shard.get_tuning()
out = {}
for var in variants:
    t = N.Tuning.from_buffer_copy(bytes(base))
    if var != "-":
        t.update(**dict(kv.split("=", 1) for kv in var.split(",")))
    shard.set_tuning(t)
    res = {}
    try:
        for b in batches:
            q = queries[b]
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
                s1, i1 = (t.clone() for t in shard.search(q, k, mode))
                s0, i0 = shard.search(q, k, "verify")
                torch.cuda.synchronize()
                a, r = i1.cpu().tolist(), i0.cpu().tolist()
                res[b]["recall"] = round(sum(len(set(x) & set(y)) for x, y in zip(a, r)) / (len(r) * k), 5)
                res[b]["max_rel_err"] = float(((s1 - s0).abs() / s0.abs().clamp_min(1e-3)).max())
            res[b]["family"] = shard.plan(b, k, mode)[0]
    except Exception as exc:  # noqa: BLE001 - report and go on to the next variant
        res["error"] = f"{type(exc).__name__}: {exc}"[:300]
    out[var] = res
    if len(variants) > 1:
        print(var, "->", json.dumps(res), flush=True)
print(json.dumps(out if len(variants) > 1 else out[variants[0]]))
