"""GPU bring-up: run each kernel family against the oracle and print compact diagnostics.
Usage (on a GPU box):  python tools/gpu_bringup.py [stream|tensor|pool|perf|all]
Each section is independent; failures are reported, not raised, so one call covers everything.
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (checker only)
import vietnamese_qa_system_b200 as vqa  # noqa: E402
from vietnamese_qa_system_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)


def mk(n, d, b, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    docs = torch.randn(n, d, generator=g)
    q = torch.randn(b, d, generator=g)
    docs = torch.from_numpy(oracle.normalize_rows(docs.numpy()))
    q = torch.from_numpy(oracle.normalize_rows(q.numpy()))
    return docs.to(dtype), q


def check(name, docs, q, k, mode, storage):
    rows = docs.to(DEV)
    shard = ops.FlatShard(rows)
    s, i = shard.search(q.to(DEV), k, mode)
    torch.cuda.synchronize()
    s, i = s.cpu().numpy(), i.cpu().numpy()
    os_, oi = oracle.search(docs.float().numpy(), q.numpy(), k, oracle.CANONICAL, storage)
    ids_ok = np.array_equal(i, oi)
    bits_ok = np.array_equal(s.view(np.int32), os_.view(np.int32))
    fin = np.isfinite(os_)
    err = np.abs(s[fin] - os_[fin]).max() if fin.any() else 0.0
    rec = np.mean([len(set(i[b]) & set(oi[b])) / max(1, (oi[b] >= 0).sum()) for b in range(i.shape[0])])
    print(f"[{name}] n={docs.shape[0]} d={docs.shape[1]} B={q.shape[0]} k={k} mode={mode}: ids_exact={ids_ok} "
          f"score_bits={bits_ok} max_abs_err={err:.3e} recall={rec:.4f}", flush=True)
    if not ids_ok:
        bad = np.argwhere(i != oi)[:3]
        for b, j in bad:
            print(f"    q{b} rank{j}: got ({i[b, j]}, {s[b, j]:.7f}) want ({oi[b, j]}, {os_[b, j]:.7f})")
    return ids_ok


def section_stream():
    for (n, d, b, k) in [(1000, 768, 4, 5), (10000, 768, 64, 5), (5000, 384, 3, 10), (777, 1024, 8, 10),
                         (3000, 200, 2, 3), (50, 768, 1, 64), (4097, 768, 1, 100)]:
        docs, q = mk(n, d, b, seed=n)
        check("stream/f32", docs, q, k, "verify", "fp32")
    for dt, st in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
        for (n, d, b, k) in [(10000, 768, 8, 10), (3001, 1024, 5, 10), (2000, 384, 2, 5), (999, 72, 3, 7)]:
            docs, q = mk(n, d, b, seed=n + 1, dtype=dt)
            check(f"stream/{st}", docs, q, k, "verify", st)


def section_tensor():
    for dt, st in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
        for (n, d, b, k) in [(128, 64, 8, 4), (1000, 768, 8, 10), (10000, 768, 32, 10), (33333, 768, 16, 10),
                             (5000, 1024, 64, 10), (20000, 768, 100, 10), (4000, 768, 32, 100)]:
            docs, q = mk(n, d, b, seed=n + 2, dtype=dt)
            try:
                check(f"tensor/{st}", docs, q, k, "tensor", st)
            except Exception as e:  # noqa: BLE001
                print(f"[tensor/{st}] n={n} d={d} B={b} k={k}: EXC {type(e).__name__}: {e}", flush=True)
                raise


def section_pool():
    g = torch.Generator().manual_seed(5)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        for (b, s, d) in [(4, 16, 768), (3, 37, 384), (2, 5, 1024), (256, 64, 768)]:
            h = torch.randn(b, s, d, generator=g).to(dt)
            lens = torch.randint(0, s + 1, (b,), generator=g)
            mask = (torch.arange(s)[None, :] < lens[:, None]).to(torch.int64)
            out = ops.pool_normalize(h.to(DEV), mask.to(DEV)).cpu().numpy()
            ref = oracle.mean_pool(h.float().numpy(), mask.numpy(), True)
            print(f"[pool] {dt} {b}x{s}x{d}: max_abs_err={np.abs(out - ref).max():.3e}", flush=True)
    x = torch.randn(1000, 768, generator=g)
    x[17] = 0
    out = ops.normalize_rows(x.to(DEV)).cpu().numpy()
    print(f"[normalize] max_abs_err={np.abs(out - oracle.normalize_rows(x.numpy())).max():.3e}")
    cs = torch.randn(4, 6, 10, generator=g)
    ci = torch.randint(0, 100000, (4, 6, 10), generator=g)
    ci[0, 0, 9] = -1
    ms, mi = ops.merge_topk(cs.to(DEV), ci.to(DEV), 10)
    rs, ri = oracle.merge_topk(cs.numpy(), ci.numpy(), 10)
    print(f"[merge] ids_exact={np.array_equal(mi.cpu().numpy(), ri)} scores={np.array_equal(ms.cpu().numpy(), rs)}")


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def section_perf(n=1_000_000, d=768):
    g = torch.Generator(device=DEV).manual_seed(1234)
    rows = torch.empty(n, d, dtype=torch.bfloat16, device=DEV)
    for lo in range(0, n, 250_000):
        blk = torch.randn(min(250_000, n - lo), d, generator=g, device=DEV)
        rows[lo:lo + blk.shape[0]] = ops.normalize_rows(blk, cast_dtype=torch.bfloat16)
    shard = ops.FlatShard(rows)
    gb = n * d * 2 / 1e9
    for b in (1, 2, 4, 8, 16, 32, 64):
        q = ops.normalize_rows(torch.randn(b, d, generator=g, device=DEV))
        for mode in ("stream", "tensor"):
            if mode == "stream" and b > 16:
                continue
            try:
                ms = timeit(lambda: shard.search(q, 10, mode))
                fam, nl = shard.plan(b, 10, mode)
                print(f"[perf] n={n} B={b} {mode}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s-equivalent "
                      f"(launches={nl}) qps={b / ms * 1e3:.0f}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"[perf] B={b} {mode}: EXC {type(e).__name__}: {e}", flush=True)
    q = ops.normalize_rows(torch.randn(32, d, generator=g, device=DEV))
    ms = timeit(lambda: torch.topk(q.to(torch.bfloat16) @ rows.T, 10, dim=1))
    print(f"[perf] torch matmul+topk B=32: {ms:.3f} ms")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), "lib", vqa._native.lib().vqa_version(), flush=True)
    secs = {"stream": section_stream, "pool": section_pool, "tensor": section_tensor, "perf": section_perf}
    for name, fn in secs.items():
        if what in (name, "all"):
            t0 = time.time()
            try:
                fn()
            except Exception:  # noqa: BLE001
                traceback.print_exc()
            print(f"--- {name} done in {time.time() - t0:.1f}s", flush=True)
