"""What paces the TMEM-resident-query kernel (ts_topk_kernel)?  Per-CTA cycle counters (vqa_debug_timeline): cycles the
TMA producer waited for a free ring stage, the MMA warp waited for documents / for a free accumulator stage, and
epilogue warp 0 waited for an accumulator -- each as a share of that warp's whole loop, median over the CTAs.
usage (on a B200): ROWS=10000000 BATCHES=128,256 [KNOBS=ts_ks=4] python tools/ts_waits_probe.py"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200 import _native as N, ops  # noqa: E402

n, d = int(os.environ.get("ROWS", "10000000")), int(os.environ.get("DIM", "768"))
k = int(os.environ.get("K", "10"))
MODE = os.environ.get("MODE", "ts")
dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[os.environ.get("DTYPE", "bf16")]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
rows = torch.empty((n, d), dtype=dt, device=dev)
for lo in range(0, n, 500000):
    m = min(500000, n - lo)
    rows[lo:lo + m] = ops.normalize_rows(torch.randn((m, d), generator=g, device=dev)).to(dt)
shard = ops.FlatShard(rows)
knobs = dict(kv.split("=", 1) for kv in os.environ.get("KNOBS", "").split(",") if kv)
if knobs:
    shard.set_tuning(**knobs)
stamps = torch.zeros((148, 32), dtype=torch.int64, device=dev)
out = {"rows": n, "dim": d, "k": k, "knobs": knobs}
for b in [int(x) for x in os.environ.get("BATCHES", "128,256").split(",")]:
    q = ops.normalize_rows(torch.randn((b, d), generator=g, device=dev))
    for _ in range(3):
        shard.search(q, k, MODE)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        shard.search(q, k, MODE)
    e1.record()
    torch.cuda.synchronize()
    N.check(N.lib().vqa_debug_timeline(shard._h, ctypes.c_void_p(stamps.data_ptr()), stamps.numel() * 8))
    stamps.zero_()
    shard.search(q[:256] if b > 256 else q, k, MODE)        # one scan launch
    torch.cuda.synchronize()
    N.check(N.lib().vqa_debug_timeline(shard._h, None, 0))
    t = stamps.cpu().numpy().astype(np.float64)
    live = t[:, 8] > 0
    med = lambda x: float(np.median(x[live]))  # noqa: E731
    mma_loop, epi_loop, prod_loop = med(t[:, 4]), med(t[:, 6]), med(t[:, 9])
    tiles = med(t[:, 8])
    out[f"b{b}"] = {
        "search_ms": round(e0.elapsed_time(e1) / 5, 4), "ctas": int(live.sum()), "tiles_per_cta": tiles,
        "kernel_us": round(float((t[live, 15].max() - t[live, 0].min()) / 1e3), 1),
        "mma_loop_cycles_per_tile": round(mma_loop / tiles), "mma_floor_cycles_per_tile": (d // 16) * 32,
        "mma_wait_documents_share": round(med(t[:, 2]) / mma_loop, 3),
        "mma_wait_accumulator_share": round(med(t[:, 3]) / mma_loop, 3),
        "producer_wait_free_stage_share": round(med(t[:, 1]) / max(prod_loop, 1.0), 3),
        "epilogue_wait_accumulator_share": round(med(t[:, 5]) / epi_loop, 3),
        "epilogue_busy_cycles_per_tile": round((epi_loop - med(t[:, 5])) / tiles),
        "epilogue_flushes_per_cta": med(t[:, 7]),
        # where the rest of a launch goes: per-CTA life time (globaltimer) against its role loops (SM cycles)
        "cta_us_min_med_max": [round(float(x) / 1e3, 1) for x in np.percentile((t[live, 15] - t[live, 0]), [0, 50, 100])],
        "cta_start_spread_us": round(float(t[live, 0].max() - t[live, 0].min()) / 1e3, 1),
        "cta_end_spread_us": round(float(t[live, 15].max() - t[live, 15].min()) / 1e3, 1),
        "loop_cycles_mma_epilogue_producer": [round(mma_loop), round(epi_loop), round(prod_loop)]}
print(json.dumps(out))
