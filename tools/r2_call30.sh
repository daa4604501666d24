#!/bin/bash
# Round-2 GPU call 30 (one B200): k > 128 with half as many segments (k / 64): tests and
# timings at the 10 M x 768 index.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== wide-k tests"
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_sparse.py -m gpu -q --tb=short -k "wide or merge_segments or error_behaviour or hybrid" 2>&1 | tail -n 40
echo "== timings"
timeout 600 python - <<'PY'
import torch, time
from vietnamese_qa_system_b200 import ops
dev = torch.device("cuda", 0)
n, d = 10_000_000, 768
rows = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
for a in range(0, n, 1_000_000):
    x = torch.randn((1_000_000, d), generator=g, device=dev)
    rows[a:a + 1_000_000] = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
shard = ops.FlatShard(rows)
for b in (1, 8, 32):
    q = torch.randn((b, d), generator=g, device=dev); q = q / q.norm(dim=1, keepdim=True)
    for k in (100, 128, 130, 300, 1000):
        for _ in range(2):
            shard.search(q, k, "fast")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            s, i = shard.search(q, k, "fast")
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"B={b} k={k}: {ms:.2f} ms per search ({len(shard._segs)} cached segments)", flush=True)
PY
echo "== done"
