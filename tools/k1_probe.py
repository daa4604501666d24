"""K1 (vqa_pool_normalize) probe: runs the kernel on three shapes (for ncu), and with an argument also times the host
enqueue cost per call against the drained time.  usage: python tools/k1_probe.py [host]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
shapes = ((256, 256, 768, torch.bfloat16), (32, 64, 768, torch.bfloat16), (256, 256, 768, torch.float32))
for hb, hs, dim, dt in shapes:
    hidden = torch.randn((hb, hs, dim), generator=g, device=dev).to(dt)
    lens = torch.randint(16, hs + 1, (hb,), generator=g, device=dev)
    mask = (torch.arange(hs, device=dev)[None, :] < lens[:, None]).to(torch.int64)
    copies = [hidden] + [hidden.clone() for _ in range(3)]
    for i in range(8):
        ops.pool_normalize(copies[i & 3], mask)
    torch.cuda.synchronize()
    if len(sys.argv) > 1:
        t0 = time.perf_counter()
        for i in range(2000):
            ops.pool_normalize(copies[i & 3], mask)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"[{hb},{hs},{dim}] {dt}: host enqueue {1e6 * (t1 - t0) / 2000:.1f} us per call, drained after "
              f"{1e6 * (t2 - t0) / 2000:.1f} us per call", flush=True)
