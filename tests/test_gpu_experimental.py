"""GPU parity tests of the kernel VARIANTS behind the tuning knobs (include/vqa.h vqa_tuning_t), each against the
CPU oracle and against the variant it replaces:
  * reduce_select -- radix-select candidate reduce for k > 32 and CTA-per-query re-scoring (DEFAULT since round 2)
  * ts_qs [ts_ks=n] -- TMEM-resident-query kernel with part of the query block in shared memory: dim <= 1024, more
    accumulator stages at dim 768 (DEFAULT since round 2)
  * reduce_early -- early exit in the k <= 32 candidate reduce
  * pdl_chain -- consecutive scan launches of one search overlap the previous reduce (PDL without a wait)
Round 1 gated this file behind VQA_EXPERIMENTAL=1 because none of it had met the hardware; the driver's round-1
bench ran all of it cleanly on a B200, so the gate is gone.  Knobs are given as VQA_* variables where a test
builds a fresh index per call (they are parsed once, in vqa_index_create) and through FlatShard.set_tuning where
one index is searched under several settings.  Same bars as tests/test_gpu_search.py."""
import numpy as np
import pytest
import torch

import oracle
from tests.conftest import unit_rows
from tests.test_gpu_search import gpu_search, recall

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check(docs, q, k, mode, storage):
    s, i, stored = gpu_search(docs, q, k, mode, storage)
    os_, oi = oracle.search(stored, q, k, oracle.CANONICAL, storage)
    assert recall(i, oi) >= 0.999
    fin = np.isfinite(os_)
    assert np.all(np.abs(s[fin] - os_[fin]) <= 1e-5 * np.abs(os_[fin]) + 2e-6)
    assert np.all(np.diff(s[:, :min(k, docs.shape[0])], axis=1) <= 0)
    return s, i


@pytest.mark.parametrize("storage", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("n,d,b,k", [(4000, 768, 32, 100), (130, 768, 5, 128), (50000, 384, 9, 33), (20, 64, 3, 40)])
def test_radix_select_reduce_is_bit_identical_to_the_list_insertion_reduce(monkeypatch, storage, n, d, b, k):
    """k > 32 in every kernel family: same ids and score bits with either reduce kernel BEHIND THE SAME SCAN; verify
    mode stays bit-identical to the canonical oracle.  (ts_qs=0 pins the scan: with the QS kernel, fp16 rows and
    reduce_select the planner switches to the big-k screen + exact re-score, whose scores differ by design --
    that route is checked against the oracle in test_query_block_split_between_tmem_and_smem.)"""
    monkeypatch.setenv("VQA_TS_QS", "0")
    rng = np.random.default_rng(n + k)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[1]                                     # a tie, decided by the lower id
    for mode in ("verify", "fast") if storage != "fp32" else ("verify",):
        out = []
        for sel in ("0", "1"):
            monkeypatch.setenv("VQA_REDUCE_SELECT", sel)
            s, i, stored = gpu_search(docs, q, k, mode, storage)
            out.append((s, i))
        assert np.array_equal(out[0][1], out[1][1])
        assert np.array_equal(out[0][0].view(np.int32), out[1][0].view(np.int32))
        if mode == "verify":
            os_, oi = oracle.search(stored, q, k, oracle.CANONICAL, storage)
            assert np.array_equal(out[1][1], oi) and np.array_equal(out[1][0].view(np.int32), os_.view(np.int32))


@pytest.mark.parametrize("storage,n,d,b,k,ks,select", [
    ("fp16", 30000, 1024, 64, 100, 4, 1),    # BASELINE configs[3] in miniature: big-k screen + radix-select re-score
    ("fp16", 30000, 1024, 64, 100, 8, 1),
    ("bf16", 30000, 1024, 64, 100, 4, 1),    # bf16 keeps hi/lo rows for big k
    ("bf16", 20000, 1024, 64, 10, 4, 0),     # dim 1024 top-10: register lists, warp reduce re-scores 32
    ("bf16", 20000, 1024, 300, 10, 4, 0),    # three chunks of 128: cluster of 4 with TMA multicast
    ("fp16", 5000, 832, 40, 5, 1, 0),        # 13 query blocks
    ("bf16", 20000, 768, 100, 10, 4, 0),     # dim 768, 4 accumulator stages
    ("bf16", 20000, 768, 256, 10, 6, 0),     # 5 accumulator stages, cluster of 2
    ("fp16", 20000, 768, 130, 10, 10, 0),    # two blocks in TMEM, ten in shared memory, cluster of 2
    ("bf16", 4000, 768, 32, 100, 2, 1),      # k = 100 at dim 768 with the QS variant (hi/lo heaps)
])
def test_query_block_split_between_tmem_and_smem(monkeypatch, storage, n, d, b, k, ks, select):
    monkeypatch.setenv("VQA_TS_QS", "1")
    monkeypatch.setenv("VQA_TS_KS", str(ks))
    monkeypatch.setenv("VQA_REDUCE_SELECT", str(select))
    rng = np.random.default_rng(n + b + ks)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    s, i = _check(docs, q, k, "ts", storage)
    assert i[0, :2].tolist() == [3, n // 2]
    s2, i2 = _check(docs, q, k, "fast", storage)               # FAST routes the same shapes to the same kernel
    assert np.array_equal(i, i2) and np.array_equal(s, s2)


def test_full_size_config_d_shard_properties(monkeypatch):
    """2 M x 1024 fp16, B = 64, top-100 on the QS path: planted duplicates come back lower id first, results are
    idempotent and independent of the batch they were asked in, and agree with the default kernel family."""
    from vietnamese_qa_system_b200 import ops

    n, d, b, k = 2_000_000, 1024, 64, 100
    g = torch.Generator(device=DEV).manual_seed(3)
    rows = torch.empty((n, d), dtype=torch.float16, device=DEV)
    for lo in range(0, n, 500_000):
        rows[lo:lo + 500_000] = ops.normalize_rows(torch.randn((500_000, d), generator=g, device=DEV)).half()
    rows[1_500_000] = rows[17]
    q = ops.normalize_rows(torch.randn((b, d), generator=g, device=DEV))
    q[0] = rows[17].float()
    shard = ops.FlatShard(rows)
    shard.set_tuning(ts_qs=0, reduce_select=0)                 # round-1 routing: smem-resident kernel, list-insert reduce
    s_ref, i_ref = (t.clone() for t in shard.search(q, k, "fast"))
    shard.set_tuning(ts_qs=1, reduce_select=1)                 # round-2 default: QS kernel, radix-select re-score
    assert shard.plan(b, k, "fast")[0] == 4
    s1, i1 = shard.search(q, k, "fast")
    s2, i2 = shard.search(q, k, "fast")
    s3, i3 = shard.search(q[:7], k, "fast")
    torch.cuda.synchronize()
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    assert torch.equal(i1[:7], i3) and torch.equal(s1[:7], s3)
    assert i1[0, :2].tolist() == [17, 1_500_000]
    a, r = i1.cpu().tolist(), i_ref.cpu().tolist()
    assert sum(len(set(x) & set(y)) for x, y in zip(a, r)) / (b * k) >= 0.999
    assert float(((s1 - s_ref).abs() / s_ref.abs()).max()) < 1e-5


@pytest.mark.parametrize("storage,mode,n,d,b,k", [("bf16", "tensor", 300000, 768, 32, 10), ("fp16", "ts", 100000, 768, 100, 10),
                                                  ("fp32", "verify", 60000, 384, 7, 32), ("bf16", "stream", 5000, 200, 3, 5)])
def test_early_exit_reduce_is_exact(monkeypatch, storage, mode, n, d, b, k):
    rng = np.random.default_rng(n + b)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    monkeypatch.setenv("VQA_REDUCE_EARLY", "0")
    s0, i0, _ = gpu_search(docs, q, k, mode, storage)
    monkeypatch.setenv("VQA_REDUCE_EARLY", "1")
    s1, i1, _ = gpu_search(docs, q, k, mode, storage)
    assert np.array_equal(i0, i1) and np.array_equal(s0.view(np.int32), s1.view(np.int32))


@pytest.mark.parametrize("storage,n,d,b,k", [("bf16", 60000, 768, 128, 10), ("fp16", 50000, 384, 40, 5),
                                             ("bf16", 30000, 768, 300, 10)])
def test_screen_mode_rescoring_through_the_select_kernel(monkeypatch, storage, n, d, b, k):
    """reduce_select also takes over the k <= 32 screen-then-rescore reduce of the TMEM-resident-query
    kernel: a CTA per query, coalesced row reads, one warp per candidate.  Same bars; ids equal the warp-per-query
    reduce's, scores to fp32 rounding of a different summation order."""
    rng = np.random.default_rng(n + b)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    monkeypatch.setenv("VQA_REDUCE_SELECT", "0")
    s0, i0, _ = gpu_search(docs, q, k, "ts", storage)
    monkeypatch.setenv("VQA_REDUCE_SELECT", "1")
    s1, i1 = _check(docs, q, k, "ts", storage)
    assert recall(i1, i0) >= 0.999 and np.abs(s1 - s0).max() <= 5e-7
    assert i1[0, :2].tolist() == [3, n // 2]


@pytest.mark.parametrize("mode,b", [("ts", 600), ("tensor", 300), ("fast", 1024)])
def test_pdl_chained_scan_launches_give_the_same_results(monkeypatch, mode, b):
    """VQA_PDL_CHAIN=1: the 2nd, 3rd ... scan launch of one search is a programmatic dependent launch without a wait,
    so it overlaps the previous launch's reduce; nothing it reads is written by that reduce."""
    rng = np.random.default_rng(b)
    docs, q = unit_rows(rng, 150000, 768), unit_rows(rng, b, 768)
    monkeypatch.setenv("VQA_PDL_CHAIN", "0")
    s0, i0, _ = gpu_search(docs, q, 10, mode, "bf16")
    monkeypatch.setenv("VQA_PDL_CHAIN", "1")
    for _ in range(3):
        s1, i1, _ = gpu_search(docs, q, 10, mode, "bf16")
        assert np.array_equal(i0, i1) and np.array_equal(s0.view(np.int32), s1.view(np.int32))


@pytest.mark.parametrize("storage,n,d,b,k", [("bf16", 60000, 768, 128, 10), ("fp16", 40000, 384, 70, 5),
                                             ("bf16", 30000, 768, 300, 10), ("bf16", 4000, 768, 8, 100),
                                             ("bf16", 50000, 768, 1, 10), ("fp16", 50000, 768, 2, 10)])
def test_round1_routing_still_passes_the_oracle_bars(monkeypatch, storage, n, d, b, k):
    """The variants the round-2 defaults replaced stay reachable through the knobs and stay correct: classic
    TMEM-resident-query kernel (whole query block in tensor memory), warp-per-query re-scoring reduce, list-insertion
    reduce for k > 32, and B <= 2 on the tcgen05 kernel instead of the streaming kernel."""
    monkeypatch.setenv("VQA_TS_QS", "0")
    monkeypatch.setenv("VQA_REDUCE_SELECT", "0")
    monkeypatch.setenv("VQA_STREAM_MAX_B", "0")
    rng = np.random.default_rng(n + b)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    s, i = _check(docs, q, k, "fast", storage)
    assert i[0, :2].tolist() == [3, n // 2]


def test_set_tuning_changes_the_plan_of_a_live_index_and_rejects_nonsense():
    from vietnamese_qa_system_b200 import ops

    rng = np.random.default_rng(9)
    rows = torch.from_numpy(unit_rows(rng, 20000, 768)).to(DEV).to(torch.bfloat16)
    q = torch.from_numpy(unit_rows(rng, 1, 768)).to(DEV)
    shard = ops.FlatShard(rows)
    assert shard.plan(1, 10, "fast")[0] == 3                   # B = 1 on a small shard: tcgen05 kernel
    shard.set_tuning(stream_max_b=2, stream_min_mb=0)
    assert shard.plan(1, 10, "fast")[0] == 2                   # streaming kernel (the plan cache was dropped)
    want = [t.clone() for t in shard.search(q, 10, "fast")]
    shard.set_tuning(stream_max_b=0)
    assert shard.plan(1, 10, "fast")[0] == 3                   # back on the tcgen05 kernel
    got = shard.search(q, 10, "fast")
    torch.cuda.synchronize()
    assert torch.equal(want[1], got[1]) and float((want[0] - got[0]).abs().max()) < 2e-6
    with pytest.raises(ValueError):
        shard.set_tuning(ts_extra=-5)
    assert shard.get_tuning().ts_extra == 6                    # a refused tuning leaves the handle untouched


@pytest.mark.parametrize("storage,n,d,b,k", [("bf16", 300000, 768, 32, 10), ("bf16", 200000, 768, 1, 10),
                                             ("fp16", 150000, 384, 16, 32), ("bf16", 3000, 768, 8, 5),
                                             ("bf16", 150000, 768, 100, 10), ("bf16", 100, 768, 32, 10)])
def test_warmup_seed_and_dynamic_tiles_leave_results_unchanged(monkeypatch, storage, n, d, b, k):
    """The headline kernel's warm-up seed (a bound from the first tiles' per-warp maxima) and its dynamic tile schedule
    change WHEN candidates are seen, never which are kept: ids and score bits equal the run with both switched off,
    three times in a row on one index (slots and tile counter are recycled), and pass the oracle bars."""
    from vietnamese_qa_system_b200 import ops

    rng = np.random.default_rng(n + b)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    monkeypatch.setenv("VQA_SEED", "0")
    monkeypatch.setenv("VQA_DYN_TILES", "0")
    s0, i0, stored = gpu_search(docs, q, k, "tensor", storage)
    monkeypatch.delenv("VQA_SEED")
    monkeypatch.delenv("VQA_DYN_TILES")
    rows = torch.from_numpy(docs).to(DEV).to({"bf16": torch.bfloat16, "fp16": torch.float16}[storage])
    shard = ops.FlatShard(rows)
    t = shard.get_tuning()
    assert t.seed == 1 and t.dyn_tiles == 1
    qd = torch.from_numpy(q).to(DEV)
    for _ in range(3):
        s1, i1 = shard.search(qd, k, "tensor")
        torch.cuda.synchronize()
        assert np.array_equal(i0, i1.cpu().numpy()) and np.array_equal(s0.view(np.int32), s1.cpu().numpy().view(np.int32))
    os_, oi = oracle.search(stored, q, k, oracle.CANONICAL, storage)
    assert recall(i0, oi) >= 0.999 and i0[0, :2].tolist() == [3, n // 2]


def test_warmup_seed_under_graph_replay_with_new_queries():
    """A CUDA-graph replay reuses the captured epoch: the reduce must have cleared the seed slots and the tile counter,
    or the maxima of the PREVIOUS queries would pass as bounds for the new ones."""
    from vietnamese_qa_system_b200 import ops

    rng = np.random.default_rng(5)
    rows = torch.from_numpy(unit_rows(rng, 100000, 768)).to(DEV).to(torch.bfloat16)
    shard = ops.FlatShard(rows)
    ref = ops.FlatShard(rows)
    ref.set_tuning(seed=0, dyn_tiles=0)
    q = torch.from_numpy(unit_rows(rng, 32, 768)).to(DEV)
    out_s = torch.empty((32, 10), dtype=torch.float32, device=DEV)
    out_i = torch.empty((32, 10), dtype=torch.int64, device=DEV)
    shard.search(q, 10, "tensor", out_s, out_i)             # warm-up outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        shard.search(q, 10, "tensor", out_s, out_i)
    for rep in range(3):
        q.copy_(torch.from_numpy(unit_rows(np.random.default_rng(rep), 32, 768)).to(DEV))
        g.replay()
        torch.cuda.synchronize()
        ref_s, ref_i = ref.search(q, 10, "tensor")
        torch.cuda.synchronize()
        assert torch.equal(out_i, ref_i) and torch.equal(out_s, ref_s)


@pytest.mark.parametrize("storage,n,d,b,k", [("bf16", 60000, 768, 256, 10), ("fp16", 50000, 768, 200, 10),
                                             ("bf16", 30000, 768, 300, 5), ("bf16", 40000, 1024, 256, 10),
                                             ("fp16", 9000, 384, 129, 26), ("bf16", 200, 768, 256, 10),
                                             ("bf16", 50000, 768, 100, 10), ("fp16", 30000, 768, 64, 10),
                                             ("bf16", 20000, 1024, 33, 16)])
def test_cta_pair_kernel_matches_the_oracle(storage, n, d, b, k):
    """ts_pair_topk_kernel (tcgen05.mma.cta_group::2: M = 256 queries across a CTA pair, N = 128 documents, each CTA
    holding half of every tile; pair.cuh) + the re-scoring reduce: same bars as every fast mode, ids equal to the
    TMEM-resident-query kernel's, planted duplicate lower id first; launches of <= 128 queries (small batches, the
    tail of a large one) run the same 128-document tiles on single CTAs (PAIR = false)."""
    rng = np.random.default_rng(n + b)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    q[b - 1] = docs[n - 1]
    s, i = _check(docs, q, k, "pair", storage)
    assert i[0, :2].tolist() == [3, n // 2] and i[b - 1, 0] == n - 1
    s2, i2, _ = gpu_search(docs, q, k, "ts", storage)
    assert recall(i, i2) >= 0.999 and np.abs(s - s2).max() <= 5e-7


@pytest.mark.parametrize("storage,n,d,b,k,ks", [("fp16", 30000, 1024, 64, 100, 6), ("bf16", 20000, 768, 40, 10, 2),
                                                ("bf16", 5000, 768, 64, 26, 4), ("fp16", 9000, 512, 33, 5, 0),
                                                ("fp16", 40000, 1024, 50, 100, 8)])
def test_m64_variant_of_the_tmem_resident_query_kernel(monkeypatch, storage, n, d, b, k, ks):
    """<= 64 queries in screen mode (BASELINE configs[3]: B = 64, top-100) with M = 64 instructions: the 64 query rows
    sit in lanes 0..15 of each quarter of tensor memory, half the A bytes are read per MMA, shared-memory query blocks
    are 8 KB.  Same oracle bars; ids equal to the M = 128 kernel's."""
    monkeypatch.setenv("VQA_TS_KS", str(ks))
    monkeypatch.setenv("VQA_TS_M64", "0")
    rng = np.random.default_rng(n + b)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[3]
    q[0] = docs[3]
    q[b - 1] = docs[n - 1]
    s0, i0, _ = gpu_search(docs, q, k, "ts", storage)
    monkeypatch.setenv("VQA_TS_M64", "1")
    s, i = _check(docs, q, k, "ts", storage)
    assert i[0, :2].tolist() == [3, n // 2] and i[b - 1, 0] == n - 1
    assert recall(i, i0) >= 0.999 and np.abs(s - s0).max() <= 5e-7


@pytest.mark.parametrize("storage", ["bf16", "fp16"])
@pytest.mark.parametrize("n,d,b,k", [(30000, 768, 32, 10), (9000, 768, 17, 5), (20000, 384, 8, 26), (4000, 1024, 32, 1),
                                     (150, 768, 3, 10)])
def test_screen_mode_of_the_smem_resident_kernel(monkeypatch, storage, n, d, b, k):
    """ss_screen (planner default for scans of >= 12 GB with more than 16 queries -- BASELINE's 10 M x 768 index on ONE
    GPU at B = 32; forced here on small indexes): mma_topk_kernel with one storage-precision column per query (half the
    tensor work of the hi/lo form), k + 6 candidates per list, exact fp32 re-scoring in the reduce.  Against the oracle
    (recall, exact scores) and against the hi/lo form of the same kernel."""
    rng = np.random.default_rng(n + d + k)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    docs[n // 2] = docs[3]                                    # a tie, decided by the lower id
    q[0] = docs[3]
    monkeypatch.setenv("VQA_SS_SCREEN", "1")
    s1, i1 = _check(docs, q, k, "tensor", storage)
    monkeypatch.setenv("VQA_SS_SCREEN", "0")
    s0, i0 = _check(docs, q, k, "tensor", storage)
    assert recall(i1, i0) >= 0.999
    if k >= 2 and n > 2000:
        assert i1[0, :2].tolist() == [3, n // 2]
