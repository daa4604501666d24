"""Bring-up of the TMEM-resident-query kernel (mode "ts"): parity against the oracle + timing."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from vietnamese_qa_system_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)


def check(n, d, b, k, dt=torch.bfloat16, st="bf16"):
    g = torch.Generator().manual_seed(n + b)
    docs = torch.from_numpy(oracle.normalize_rows(torch.randn(n, d, generator=g).numpy())).to(dt)
    q = torch.from_numpy(oracle.normalize_rows(torch.randn(b, d, generator=g).numpy()))
    shard = ops.FlatShard(docs.to(DEV))
    s, i = shard.search(q.to(DEV), k, "ts")
    torch.cuda.synchronize()
    s, i = s.cpu().numpy(), i.cpu().numpy()
    os_, oi = oracle.search(docs.float().numpy(), q.numpy(), k, oracle.SEMANTIC, st)
    rec = np.mean([len(set(i[r]) & set(oi[r])) / max(1, (oi[r] >= 0).sum()) for r in range(b)])
    fin = np.isfinite(os_)
    err = np.abs(s[fin] - os_[fin]).max()
    print(f"[ts] n={n} d={d} B={b} k={k} {st} split={os.environ.get('VQA_TS_SPLIT','1')} afp16={os.environ.get('VQA_TS_AFP16','0')}: "
          f"ids_exact={np.array_equal(i, oi)} recall={rec:.4f} max_abs_err={err:.2e}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    if what == "check":
        for (n, d, b, k) in [(64, 64, 4, 3), (1000, 768, 8, 10), (10000, 768, 64, 10), (33333, 768, 100, 10),
                             (5000, 384, 200, 10), (4000, 768, 64, 100), (20000, 768, 300, 10)]:
            check(n, d, b, k)
        check(10000, 768, 64, 10, torch.float16, "fp16")
    else:
        import json
        n, d = int(os.environ.get("ROWS", "10000000")), 768
        g = torch.Generator(device=DEV).manual_seed(1)
        rows = torch.empty((n, d), dtype=torch.bfloat16, device=DEV)
        for lo in range(0, n, 500000):
            m = min(500000, n - lo)
            rows[lo:lo + m] = ops.normalize_rows(torch.randn((m, d), generator=g, device=DEV), cast_dtype=torch.bfloat16)
        shard = ops.FlatShard(rows)
        res = {}
        for b in (32, 64, 128, 256, 512):
            q = ops.normalize_rows(torch.randn((b, d), generator=g, device=DEV))
            sv, iv = shard.search(q, 10, "tensor")
            for mode in ("ts",):
                for _ in range(2):
                    s, i = shard.search(q, 10, mode)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    shard.search(q, 10, mode)
                e1.record()
                torch.cuda.synchronize()
                rec = np.mean([len(set(a.tolist()) & set(c.tolist())) / 10 for a, c in zip(i.cpu(), iv.cpu())])
                res[b] = {"ms": round(e0.elapsed_time(e1) / 5, 3), "recall_vs_tensor": round(float(rec), 4)}
        print(json.dumps(res))
