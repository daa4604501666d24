"""GPU probe of the sparse (BM25) leg at scale: synthetic Zipf postings built directly as CSR, timed with
CUDA events.  Writes gpurun_out/sparse_probe.json.  Usage: python tools/sparse_probe.py [n_docs] [avg_len]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200.scoring import BM25  # noqa: E402


def synth(n_docs, avg_len, vocab, seed=1):
    rng = np.random.default_rng(seed)
    p = np.arange(1, vocab + 1, dtype=np.float64) ** -1.1
    p /= p.sum()
    n_tok = n_docs * avg_len
    terms = rng.choice(vocab, size=n_tok, p=p).astype(np.int64)
    docs = np.repeat(np.arange(n_docs, dtype=np.int64), avg_len)
    key, freq = np.unique(terms * n_docs + docs, return_counts=True)   # sorted by (term, doc)
    t, d = key // n_docs, key % n_docs
    used, df = np.unique(t, return_counts=True)
    offsets = np.zeros(len(used) + 1, np.int64)
    np.cumsum(df, out=offsets[1:])
    lengths = np.full(n_docs, avg_len, np.int32)
    return offsets, d.astype(np.int32), freq.astype(np.int32), lengths, [f"w{int(u)}" for u in used]


def main():
    n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    avg_len = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    t0 = time.time()
    offsets, docs, freqs, lengths, vocab = synth(n_docs, avg_len, 200_000)
    bm = BM25({"method": "bm25", "terms": True, "normalize": True})
    bm.index_postings(offsets, docs, freqs, lengths, vocab)
    build_s = time.time() - t0
    df = np.diff(offsets)
    rng = np.random.default_rng(2)
    rare_ids = np.flatnonzero((df > 50) & (df <= 0.1 * n_docs))
    common_ids = np.flatnonzero(df > 0.1 * n_docs)
    out = {"n_docs": n_docs, "postings": int(len(docs)), "terms": len(vocab), "build_s": round(build_s, 2),
           "common_terms": int(len(common_ids)), "rows": []}
    profile_only = os.environ.get("SPARSE_PROBE_PROFILE") == "1"   # under ncu: one warm + one measured launch per config
    for batch in ((256,) if profile_only else (1, 32, 256)):
        for kind in ("rare", "mixed", "common-only"):
            qs = []
            for _ in range(batch):
                if kind == "common-only":
                    q = [vocab[int(i)] for i in rng.choice(common_ids, 2, replace=False)]
                else:
                    q = [vocab[int(i)] for i in rng.choice(rare_ids, 4, replace=False)]
                    if kind == "mixed":
                        q.append(vocab[int(rng.choice(common_ids))])
                qs.append(q)
            if profile_only:
                st = bm.stage_plan(*bm.plan_queries(qs, 10), 10)
                bm.launch_staged(st)
                bm.launch_staged(st)
                torch.cuda.synchronize()
                continue
            t0 = time.perf_counter()
            for _ in range(5):
                plan = bm.plan_queries(qs, 10)
            plan_ms = (time.perf_counter() - t0) / 5 * 1e3
            q_terms, _, q_meta, _ = plan
            bytes_ = 0
            for r in range(batch):
                for j in range(int(q_meta[r, 0])):                       # accumulate-everywhere terms
                    bytes_ += int(df[q_terms[r, j]]) * 8
            for _ in range(3):
                bm.search_planned(*plan, 10)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            ev0.record()
            for _ in range(reps):
                bm.search_planned(*plan, 10)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / reps
            st = bm.stage_plan(*plan, 10)                                 # kernels only: plan already on the device
            bm.launch_staged(st)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(reps):
                bm.launch_staged(st)
            ev1.record()
            torch.cuda.synchronize()
            kms = ev0.elapsed_time(ev1) / reps
            out["rows"].append({"batch": batch, "kind": kind, "kernels_ms_per_batch": round(kms, 4),
                                "kernel_posting_gbs": round(bytes_ / kms / 1e6, 2),
                                "device_ms_per_batch": round(ms, 4),
                                "host_planning_ms": round(plan_ms, 4), "device_qps": round(batch / ms * 1e3, 1),
                                "posting_bytes": bytes_, "posting_gbs": round(bytes_ / ms / 1e6, 2)})
    if profile_only:
        return
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sparse_probe.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
