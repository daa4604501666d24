"""Tuning sweep for the tcgen05 path: each configuration runs in a fresh process (the knobs are
environment variables read by libvqa_b200.so).  usage: python tools/tune_mma.py [rows]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r"""
import os, sys, json, torch
sys.path.insert(0, os.environ["VQA_ROOT"])
from vietnamese_qa_system_b200 import ops
n, d = int(os.environ["ROWS"]), 768
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
rows = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 500000):
    m = min(500000, n - lo)
    rows[lo:lo+m] = ops.normalize_rows(torch.randn((m, d), generator=g, device=dev), cast_dtype=torch.bfloat16)
shard = ops.FlatShard(rows)
res = {}
for b in [int(x) for x in os.environ.get("BATCHES", "8,32").split(",")]:
    q = ops.normalize_rows(torch.randn((b, d), generator=g, device=dev))
    mode = os.environ.get("MODE", "tensor")
    for _ in range(3): shard.search(q, 10, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): shard.search(q, 10, mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    res[b] = round(n * d * 2 / ms / 1e6, 0)
print(json.dumps(res))
"""


def run(env_extra, rows):
    env = dict(os.environ, VQA_ROOT=ROOT, ROWS=str(rows), **{k: str(v) for k, v in env_extra.items()})
    r = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True, timeout=300)
    out = r.stdout.strip().splitlines()
    return out[-1] if r.returncode == 0 and out else f"FAIL rc={r.returncode} {r.stderr[-300:]}"


if __name__ == "__main__":
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    configs = [{}]
    configs += [{"VQA_MMA_STAGES": s} for s in (2, 3, 4, 6, 8, 10)]
    configs += [{"VQA_MMA_KPS": k} for k in (2, 3, 4, 6)]
    configs += [{"VQA_TMA_L2PROMO": v} for v in (0, 1, 2)]
    configs += [{"VQA_TMA_HINT": v} for v in (0, 2)]
    configs += [{"VQA_MMA_KPS": 2, "VQA_TMA_L2PROMO": 0}, {"VQA_MMA_KPS": 3, "VQA_TMA_HINT": 0},
                {"VQA_MMA_KPS": 4, "VQA_TMA_L2PROMO": 0, "VQA_TMA_HINT": 0}]
    for c in configs:
        print(json.dumps(c), "->", run(c, rows), "GB/s by batch", flush=True)
