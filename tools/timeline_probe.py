"""Where does the per-launch fixed cost of the smem-resident tcgen05 scan kernel go?  Per-CTA %globaltimer stamps
(vqa_debug_timeline) summarised over the 148 CTAs: for every stamp the min / median / max offset from the first
CTA's entry, in microseconds.  usage (on a B200): ROWS=1250000 BATCHES=1,16,32 python tools/timeline_probe.py"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200 import _native as N, ops  # noqa: E402

NAMES = ["entry", "first_tma", "last_tma", "first_mma", "last_commit", "queries_staged", "tile0", "tile1", "tile3",
         "tile7", "tile15", "tile31", "tile63", "last_tile", "exit"]
n, d = int(os.environ.get("ROWS", "1250000")), int(os.environ.get("DIM", "768"))
k = int(os.environ.get("K", "10"))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
rows = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 500000):
    m = min(500000, n - lo)
    rows[lo:lo + m] = ops.normalize_rows(torch.randn((m, d), generator=g, device=dev)).to(torch.bfloat16)
shard = ops.FlatShard(rows)
knobs = dict(kv.split("=", 1) for kv in os.environ.get("KNOBS", "").split(",") if kv)
if knobs:
    shard.set_tuning(**knobs)
stamps = torch.zeros((148, 32), dtype=torch.int64, device=dev)
out = {"rows": n, "dim": d, "knobs": knobs}
for b in [int(x) for x in os.environ.get("BATCHES", "1,16,32").split(",")]:
    q = ops.normalize_rows(torch.randn((b, d), generator=g, device=dev))
    for _ in range(5):
        shard.search(q, k, "tensor")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        shard.search(q, k, "tensor")
    e1.record()
    torch.cuda.synchronize()
    N.check(N.lib().vqa_debug_timeline(shard._h, ctypes.c_void_p(stamps.data_ptr()), stamps.numel() * 8))
    reps = []
    for _ in range(4):
        stamps.zero_()
        shard.search(q, k, "tensor")
        torch.cuda.synchronize()
        reps.append(stamps.cpu().numpy().copy())
    N.check(N.lib().vqa_debug_timeline(shard._h, None, 0))
    t = reps[-1][:, :16].astype(np.float64)
    c = reps[-1][:, 16:].astype(np.float64)
    t0 = t[:, 0].min()
    res = {"search_ms_untimed_by_stamps": round(e0.elapsed_time(e1) / 20, 4), "stamps_us_min_med_max": {},
           "cycles_from_entry_med": {}}
    for i, name in enumerate(NAMES):
        col = t[:, i]
        ok = col > 0
        if not ok.any():
            continue
        rel = (col[ok] - t0) / 1e3
        res["stamps_us_min_med_max"][name] = [round(float(rel.min()), 2), round(float(np.median(rel)), 2),
                                              round(float(rel.max()), 2)]
        res["cycles_from_entry_med"][name] = int(np.median(c[ok, i] - c[ok, 0]))
    res["kernel_us_first_entry_to_last_exit"] = round(float((t[:, 14].max() - t0) / 1e3), 2)
    res["per_cta_stream_us_tile0_to_last_tile_min_med_max"] = [round(float(x), 2) for x in np.percentile(
        (t[:, 13] - t[:, 6]) / 1e3, [0, 50, 100])]
    res["entry_spread_us"] = round(float((t[:, 0].max() - t0) / 1e3), 2)
    res["rep_to_rep_kernel_us"] = [round(float((r[:, 14].max() - r[:, 0][r[:, 0] > 0].min()) / 1e3), 2) for r in reps]
    out[f"b{b}"] = res
print(json.dumps(out))
