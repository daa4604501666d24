"""End-to-end (host buffers) throughput of the search: synchronous calls vs two batches in flight.
Every step copies its queries from pinned host memory and its results back (both inside the timed region); the
pipelined loop only drops the per-step host wake-up (vqa_search_host_async, two buffer slots, one final sync).
usage (on a B200): ROWS=10000000 BATCH=32 STEPS=200 python tools/e2e_pipeline_probe.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200 import ops  # noqa: E402

n, d = int(os.environ.get("ROWS", "10000000")), int(os.environ.get("DIM", "768"))
b, k, steps = int(os.environ.get("BATCH", "32")), int(os.environ.get("K", "10")), int(os.environ.get("STEPS", "200"))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
rows = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 500000):
    m = min(500000, n - lo)
    rows[lo:lo + m] = ops.normalize_rows(torch.randn((m, d), generator=g, device=dev)).to(torch.bfloat16)
shard = ops.FlatShard(rows)
qs = [ops.normalize_rows(torch.randn((b, d), generator=g, device=dev)).cpu().pin_memory() for _ in range(4)]
for q in qs:
    shard.search_host(q, k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(steps):
    shard.search_host(qs[s % 4], k)
t_sync = (time.perf_counter() - t0) / steps
want = [tuple(t.clone() for t in shard.search_host(q, k)) for q in qs]
for s in range(4):
    shard.search_host_async(qs[s % 4], k, slot=s % 2)
torch.cuda.synchronize()
pending = [None, None]
ok = True
t0 = time.perf_counter()
for s in range(steps):
    slot = s % 2
    if pending[slot] is not None:          # the result of step s-2 is read before its buffers are reused
        hs, hi, ev, idx = pending[slot]
        ev.synchronize()
        ok = ok and torch.equal(hi, want[idx][1])
    hs, hi, ev = shard.search_host_async(qs[s % 4], k, slot=slot)
    pending[slot] = (hs, hi, ev, s % 4)
for p in pending:
    if p is not None:
        p[2].synchronize()
        ok = ok and torch.equal(p[1], want[p[3]][1])
t_pipe = (time.perf_counter() - t0) / steps
print(json.dumps({"rows": n, "batch": b, "k": k, "sync_ms": round(t_sync * 1e3, 4), "pipelined_ms": round(t_pipe * 1e3, 4),
                  "sync_qps": round(b / t_sync), "pipelined_qps": round(b / t_pipe), "results_equal": bool(ok)}))
