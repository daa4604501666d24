#!/bin/bash
# Round-2 GPU call 18 (one B200): K1 bulk-copy kernel variants (speculative first blocks; 1 or 2 CTAs per sequence):
# parity tests + the timing table of call 16.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== K1 tests"
timeout 600 python -m pytest tests/test_gpu_pool.py -m gpu -q --tb=short 2>&1 | tail -n 4
echo "== timing"
timeout 600 python - <<'PY'
import torch
from vietnamese_qa_system_b200 import ops
dev = torch.device("cuda", 0)
def timed(fn, n, w):
    for i in range(w): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
g = torch.Generator(device=dev).manual_seed(5)
for hb, hs, dim, dt in ((256, 256, 768, torch.bfloat16), (256, 256, 384, torch.bfloat16), (32, 64, 768, torch.bfloat16),
                        (1024, 128, 768, torch.bfloat16), (256, 256, 768, torch.float32), (256, 512, 1024, torch.float16)):
    hidden = torch.randn((hb, hs, dim), generator=g, device=dev).to(dt)
    lens = torch.randint(16, hs + 1, (hb,), generator=g, device=dev)
    mask = (torch.arange(hs, device=dev)[None, :] < lens[:, None]).to(torch.int64)
    es = hidden.element_size()
    valid = int(lens.sum().item()) * dim * es
    copies = [hidden] + [hidden.clone() for _ in range(max(1, int(400e6 // max(valid, 1))))]
    n = len(copies)
    ms = timed(lambda i: ops.pool_normalize(copies[i % n], mask), 60, 6)
    full = timed(lambda i: ops.pool_normalize(copies[i % n], torch.ones_like(mask)), 60, 6)
    print(f"[{hb},{hs},{dim}] {str(dt)[6:]}: {ms*1e3:.1f} us for {valid/1e6:.1f} MB valid = {valid/ms/1e6:.0f} GB/s "
          f"({valid/ms/1e6/6548.2:.2f} of the measured HBM peak); unmasked {full*1e3:.1f} us = "
          f"{hb*hs*dim*es/full/1e6:.0f} GB/s ({hb*hs*dim*es/full/1e6/6548.2:.2f})", flush=True)
PY
echo "== done"
