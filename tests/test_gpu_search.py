"""Parity of the CUDA search path (through the C ABI) with the CPU oracle.

fp32 verify mode: ids AND score bits identical to the canonical oracle; scores within 1e-5
relative of the fp64 semantic oracle.  bf16/fp16 fast modes (CUDA-core streaming kernel and
tcgen05 tensor-core kernel): recall@k >= 0.999 against fp32 arithmetic on the same rows."""
import numpy as np
import pytest
import torch

import oracle
from tests.conftest import unit_rows
from tests.golden import inputs
from vietnamese_qa_system_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TORCH_DT = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def gpu_search(docs_np, q_np, k, mode, storage="fp32", first_id=0):
    rows = torch.from_numpy(docs_np).to(DEV).to(TORCH_DT[storage])
    shard = ops.FlatShard(rows, first_global_id=first_id)
    s, i = shard.search(torch.from_numpy(q_np).to(DEV), k, mode)
    torch.cuda.synchronize()
    stored = rows.float().cpu().numpy()           # the embeddings actually indexed
    return s.cpu().numpy(), i.cpu().numpy(), stored


def recall(got, want):
    hits = [len(set(g[g >= 0]) & set(w[w >= 0])) / max(1, int((w >= 0).sum())) for g, w in zip(got, want)]
    return float(np.mean(hits))


def assert_bit_exact(docs, q, k, storage="fp32", first_id=0):
    s, i, stored = gpu_search(docs, q, k, "verify", storage, first_id)
    os_, oi = oracle.search(stored, q, k, oracle.CANONICAL, storage, first_id)
    assert np.array_equal(i, oi), np.argwhere(i != oi)[:4]
    assert np.array_equal(s.view(np.int32), os_.view(np.int32))
    sem = oracle.search(stored, q, k, oracle.SEMANTIC, storage, first_id)[0]
    fin = np.isfinite(sem)
    assert np.all(np.abs(s[fin] - sem[fin]) <= 1e-5 * np.maximum(np.abs(sem[fin]), 1e-3))
    return s, i


# ---- golden vectors ---------------------------------------------------------------------
def test_config_a_fp32_verify_bit_exact(golden):
    docs, q = inputs.config_a()                     # BASELINE configs[0]: 10k x 768, 64 queries, top-5
    s, i = assert_bit_exact(docs, q, 5)
    assert np.array_equal(i, golden["cfgA_ids"])
    np.testing.assert_allclose(s, golden["cfgA_scores"], rtol=1e-5, atol=1e-7)


def test_kat1_identity(golden):
    for storage in ("fp32", "bf16", "fp16"):
        for mode in ("verify", "stream"):
            s, i, _ = gpu_search(golden["kat1_docs"], golden["kat1_q"], 8, mode, storage)
            assert i.tolist() == [[3, 0, 1, 2, 4, 5, 6, 7]] and s.tolist() == [[1, 0, 0, 0, 0, 0, 0, 0]]


def test_kat2_duplicates_lower_id_first(golden):
    docs, q = inputs.kat2()
    for storage, modes in (("fp32", ("verify",)), ("bf16", ("verify", "stream", "tensor"))):
        for mode in modes:
            s, i, _ = gpu_search(docs, q, 5, mode, storage)
            assert i[0, :3].tolist() == [17, 4711, 9999], (storage, mode, i)
            assert s[0, 0] == s[0, 1] == s[0, 2]
    _, i, _ = gpu_search(docs, q, 5, "verify")
    assert np.array_equal(i, golden["kat2_ids"])


def test_kat3_exact_scores_all_storages_and_kernels(golden):
    for storage in ("fp32", "bf16", "fp16"):
        for mode in ("verify", "stream"):
            s, i, _ = gpu_search(golden["kat3_docs"], golden["kat3_q"], 5, mode, storage)
            assert s.tolist() == [[1.0, 0.5, 0.0, -0.5, -1.0]] and i.tolist() == [[4, 3, 2, 1, 0]]
    for storage in ("bf16", "fp16"):
        s, i, _ = gpu_search(golden["kat3_docs"], golden["kat3_q"], 5, "tensor", storage)
        assert s.tolist() == [[1.0, 0.5, 0.0, -0.5, -1.0]] and i.tolist() == [[4, 3, 2, 1, 0]]


def test_kat4_k_equals_n_and_padding(golden):
    s, i, _ = gpu_search(golden["kat4_docs"], golden["kat4_q"], 37, "verify")
    assert np.array_equal(i, golden["kat4_ids_kN"])
    s, i, _ = gpu_search(golden["kat4_docs"], golden["kat4_q"], 64, "verify")
    assert np.array_equal(i, golden["kat4_ids_k64"])
    assert np.all(i[:, 37:] == -1) and np.all(np.isneginf(s[:, 37:]))
    s, i, _ = gpu_search(golden["kat4_docs"], golden["kat4_q"], 64, "tensor", "bf16")
    assert np.all(i[:, 37:] == -1) and np.all(np.isneginf(s[:, 37:])) and np.all(i[:, :37] >= 0)


def test_kat5_shard_boundary_ties_merge_on_device(golden):
    docs, q = golden["kat5_docs"], golden["kat5_q"]
    full_s, full_i, _ = gpu_search(docs, q, 10, "verify")
    assert np.array_equal(full_i, golden["kat5_ids"])
    for g in (2, 4, 8):
        per = 1024 // g
        parts = [gpu_search(docs[r * per:(r + 1) * per], q, 10, "verify", first_id=r * per)[:2] for r in range(g)]
        cs = torch.from_numpy(np.stack([p[0] for p in parts])).to(DEV)
        ci = torch.from_numpy(np.stack([p[1] for p in parts])).to(DEV)
        ms, mi = ops.merge_topk(cs, ci, 10)
        assert np.array_equal(mi.cpu().numpy(), full_i) and np.array_equal(ms.cpu().numpy(), full_s)


# ---- seeded sweeps ------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,b,k", [(1, 768, 1, 1), (31, 768, 2, 5), (1000, 768, 4, 5), (4097, 768, 1, 100),
                                     (5000, 384, 3, 10), (777, 1024, 8, 10), (3000, 200, 2, 3), (50, 768, 9, 64),
                                     (20000, 768, 17, 10), (2048, 4, 5, 7), (300, 2048, 3, 128)])
def test_fp32_verify_bit_exact(n, d, b, k):
    rng = np.random.default_rng(n * 7 + d)
    assert_bit_exact(unit_rows(rng, n, d), unit_rows(rng, b, d), k)


@pytest.mark.parametrize("storage", ["bf16", "fp16"])
@pytest.mark.parametrize("n,d,b,k", [(10000, 768, 8, 10), (3001, 1024, 5, 10), (2000, 384, 2, 5), (999, 72, 3, 7),
                                     (129, 8, 1, 3)])
def test_16bit_storage_verify_bit_exact(storage, n, d, b, k):
    rng = np.random.default_rng(n + d)
    assert_bit_exact(unit_rows(rng, n, d), unit_rows(rng, b, d), k, storage)


def test_first_global_id_offsets_ids():
    rng = np.random.default_rng(9)
    assert_bit_exact(unit_rows(rng, 700, 768), unit_rows(rng, 3, 768), 10, first_id=8_750_000)


@pytest.mark.parametrize("storage", ["bf16", "fp16"])
@pytest.mark.parametrize("mode", ["stream", "tensor", "ts", "fast"])
@pytest.mark.parametrize("n,d,b,k", [(128, 64, 8, 4), (1000, 768, 1, 10), (10000, 768, 32, 10),
                                     (33333, 768, 16, 10), (5000, 1024, 64, 10), (20000, 768, 100, 10),
                                     (4000, 768, 32, 100), (130, 768, 5, 128)])
def test_fast_modes_recall_vs_fp32_arithmetic(storage, mode, n, d, b, k):
    rng = np.random.default_rng(n + b)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    s, i, stored = gpu_search(docs, q, k, mode, storage)
    os_, oi = oracle.search(stored, q, k, oracle.CANONICAL, storage)
    assert recall(i, oi) >= 0.999
    fin = np.isfinite(os_)
    # scores: same candidates up to near-tie swaps; compare rank-wise within 1e-5 relative (+ tiny abs)
    assert np.all(np.abs(s[fin] - os_[fin]) <= 1e-5 * np.abs(os_[fin]) + 2e-6)
    assert np.all(np.diff(s[:, :min(k, n)], axis=1) <= 0)          # descending


def test_clustered_queries_recall_bf16_tensor():
    """Queries = doc + noise so true neighbours exist."""
    rng = np.random.default_rng(99)
    docs = unit_rows(rng, 50000, 768)
    picks = rng.integers(0, 50000, 64)
    q = docs[picks] + 0.1 * rng.standard_normal((64, 768)).astype(np.float32) / np.sqrt(768)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    s, i, stored = gpu_search(docs, q.astype(np.float32), 10, "tensor", "bf16")
    _, oi = oracle.search(stored, q.astype(np.float32), 10, oracle.SEMANTIC, "bf16")
    assert recall(i, oi) >= 0.999 and np.array_equal(i[:, 0], picks)


def test_family_selection_and_launch_count():
    rows = torch.zeros((4096, 768), dtype=torch.bfloat16, device=DEV)
    shard = ops.FlatShard(rows)
    assert shard.plan(1, 10, "fast")[0] == 3 and shard.plan(32, 10, "fast")[0] == 3   # 16-bit rows: tcgen05
    assert shard.plan(1, 10, "stream")[0] == 2
    odd = ops.FlatShard(torch.zeros((64, 72), dtype=torch.bfloat16, device=DEV))
    assert odd.plan(8, 10, "fast")[0] == 2                                            # dim % 64 != 0: streaming
    assert shard.plan(32, 10, "verify")[0] == 2
    assert shard.plan(32, 10, "fast")[1] == 2            # one scan launch + one reduce launch
    assert shard.plan(128, 10, "fast") == (5, 2)         # 33..128 queries: 128-document tiles on single CTAs, one pass
    assert shard.plan(256, 10, "fast") == (5, 2)         # > 128: CTA pairs (cta_group::2), one pass per 256 queries
    assert shard.plan(300, 10, "fast") == (5, 4) and shard.plan(128, 10, "ts") == (4, 2)
    assert shard.plan(128, 100, "fast")[0] == 4          # k > 32: same kernel with hi/lo rows and heap lists
    assert shard.plan(8, 100, "fast")[0] == 4 and shard.plan(8, 32, "fast")[0] == 3
    wide = ops.FlatShard(torch.zeros((256, 1024), dtype=torch.float16, device=DEV))
    assert wide.plan(128, 10, "fast")[0] == 4            # dim 1024: 10 query blocks in tensor memory + 6 in shared memory
    assert ops.FlatShard(torch.zeros((64, 1088), dtype=torch.float16, device=DEV)).plan(128, 10, "fast")[0] == 3
    rows32 = torch.zeros((4096, 768), dtype=torch.float32, device=DEV)
    assert ops.FlatShard(rows32).plan(32, 10, "fast")[0] == 2                         # fp32 rows never use tcgen05
    with pytest.raises(NotImplementedError):
        ops.FlatShard(rows32).search(torch.zeros((2, 768), device=DEV), 5, "tensor")
    with pytest.raises(NotImplementedError):
        ops.FlatShard(rows32).search(torch.zeros((2, 768), device=DEV), 5, "ts")


def test_error_behaviour():
    rows = torch.zeros((64, 768), dtype=torch.bfloat16, device=DEV)
    shard = ops.FlatShard(rows)
    q = torch.zeros((2, 768), device=DEV)
    with pytest.raises(ValueError):
        shard.search(q, 0)
    with pytest.raises(ValueError):
        shard.search(q, 1025)                      # 129..1024 are composed from segment searches (below)
    with pytest.raises(ValueError):
        shard.search(q, 200, workspace=torch.zeros(1 << 20, dtype=torch.uint8, device=DEV))
    with pytest.raises(ValueError):
        shard.search(torch.zeros((2, 384), device=DEV), 5)
    with pytest.raises(ValueError):
        shard.search(q.cpu(), 5)
    with pytest.raises(ValueError):
        ops.FlatShard(torch.zeros((4, 7), dtype=torch.bfloat16, device=DEV))
    empty = ops.FlatShard(torch.zeros((0, 768), dtype=torch.bfloat16, device=DEV))
    s, i = empty.search(q, 3)
    assert np.all(i.cpu().numpy() == -1) and np.all(np.isneginf(s.cpu().numpy()))


def test_search_host_matches_device_call():
    rng = np.random.default_rng(5)
    docs, q = unit_rows(rng, 5000, 768), unit_rows(rng, 16, 768)
    rows = torch.from_numpy(docs).to(DEV).to(torch.bfloat16)
    shard = ops.FlatShard(rows)
    s, i = shard.search(torch.from_numpy(q).to(DEV), 10, "fast")
    hs, hi = shard.search_host(torch.from_numpy(q).pin_memory(), 10, "fast")
    assert np.array_equal(hi.numpy(), i.cpu().numpy()) and np.array_equal(hs.numpy(), s.cpu().numpy())


def test_search_host_async_two_batches_in_flight():
    """vqa_search_host_async: same copies and launches without the final sync; two calls in flight on one stream
    with their own buffer slots, results valid once each call's event has completed."""
    rng = np.random.default_rng(15)
    docs = unit_rows(rng, 30000, 768)
    shard = ops.FlatShard(torch.from_numpy(docs).to(DEV).to(torch.bfloat16))
    qa, qb = (torch.from_numpy(unit_rows(rng, 32, 768)).pin_memory() for _ in range(2))
    want = [tuple(t.clone() for t in shard.search_host(q, 10, "fast")) for q in (qa, qb)]
    sa, ia, ea = shard.search_host_async(qa, 10, "fast", slot=0)
    sb, ib, eb = shard.search_host_async(qb, 10, "fast", slot=1)
    ea.synchronize()
    assert torch.equal(ia, want[0][1]) and torch.equal(sa, want[0][0])
    eb.synchronize()
    assert torch.equal(ib, want[1][1]) and torch.equal(sb, want[1][0])
    with pytest.raises(ValueError):
        shard.search_host_async(torch.from_numpy(unit_rows(rng, 4, 768)), 10)      # not pinned


def test_cuda_graph_capture_of_search():
    rng = np.random.default_rng(6)
    docs, q = unit_rows(rng, 20000, 768), unit_rows(rng, 32, 768)
    shard = ops.FlatShard(torch.from_numpy(docs).to(DEV).to(torch.bfloat16))
    qd = torch.from_numpy(q).to(DEV)
    s_ref, i_ref = shard.search(qd, 10, "fast")
    out_s, out_i = torch.empty_like(s_ref), torch.empty_like(i_ref)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        shard.search(qd, 10, "fast", out_s, out_i)
    out_s.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out_i, i_ref) and torch.equal(out_s, s_ref)


@pytest.mark.parametrize("b,k,mode", [(32, 10, "fast"), (1, 10, "fast"), (100, 10, "fast"), (300, 10, "fast"), (9, 100, "fast"),
                                      (5, 10, "verify")])
def test_two_stream_search_keeps_scans_back_to_back(b, k, mode):
    """vqa_search_2s: scan on the current stream, candidate reduce on a second stream behind the library's hand-over
    event.  Six searches in flight over two workspace slots (the serving loop of ShardedFlat.search_pipelined): every
    result equals the one-stream answer for its queries."""
    rng = np.random.default_rng(b + k)
    docs = unit_rows(rng, 60000, 768)
    shard = ops.FlatShard(torch.from_numpy(docs).to(DEV).to(torch.bfloat16))
    qs = [torch.from_numpy(unit_rows(rng, b, 768)).to(DEV) for _ in range(3)]
    want = [tuple(t.clone() for t in shard.search(q, k, mode)) for q in qs]
    side = torch.cuda.Stream(device=DEV)
    main = torch.cuda.current_stream()
    slots = [(torch.zeros(shard.workspace_bytes(b, k, mode), dtype=torch.uint8, device=DEV),
              torch.empty((b, k), dtype=torch.float32, device=DEV), torch.empty((b, k), dtype=torch.int64, device=DEV))
             for _ in range(2)]
    done = [None, None]
    got = []
    for step in range(6):
        sl = step % 2
        ws, out_s, out_i = slots[sl]
        if done[sl] is not None:
            done[sl][0].synchronize()
            got.append((done[sl][1], out_s.clone(), out_i.clone()))
            main.wait_event(done[sl][0])
        shard.search(qs[step % 3], k, mode, out_s, out_i, workspace=ws, reduce_stream=side)
        ev = torch.cuda.Event()
        ev.record(side)
        done[sl] = (ev, step % 3)
    for sl in range(2):
        done[sl][0].synchronize()
        got.append((done[sl][1], slots[sl][1].clone(), slots[sl][2].clone()))
    assert len(got) == 6
    for j, s, i in got:
        assert torch.equal(i, want[j][1]) and torch.equal(s, want[j][0])
    with pytest.raises(ValueError):
        shard.search(qs[0], k, mode, reduce_stream=main)         # the two streams must differ


# ---- k > 128: segment searches + vqa_merge_segments -----------------------------------------
@pytest.mark.parametrize("storage,mode", [("fp32", "verify"), ("bf16", "verify"), ("bf16", "fast"), ("fp16", "fast")])
@pytest.mark.parametrize("n,b,k", [(6000, 3, 129), (20000, 5, 300), (50000, 2, 1000), (900, 4, 1000), (131, 1, 130)])
def test_wide_k_is_composed_from_segment_searches(storage, mode, n, b, k):
    """txtai's hybrid search asks the dense leg for 10 x limit candidates (heavy_ranker.py:98,100 with limit > 12
    -> k > 128).  One vqa_search call answers k <= 128; ops.FlatShard cuts the shard into row segments, takes each
    segment's best 128 and merges them (vqa_merge_segments), halving segments the merge reports as saturated.  A
    planted run of 400 near-copies of query 0 puts most of its answer into one or two segments, so the second
    pass is exercised; exact ties keep the lower id first.  verify mode: ids and score bits equal the oracle's."""
    d = 256
    rng = np.random.default_rng(n + k)
    docs, q = unit_rows(rng, n, d), unit_rows(rng, b, d)
    if n >= 6000:
        near = q[0] + 0.02 * rng.standard_normal((400, d)).astype(np.float32)
        docs[n // 2:n // 2 + 400] = near / np.linalg.norm(near, axis=1, keepdims=True)
        docs[n // 2 + 3] = docs[n // 2 + 1]
        docs[n - 1] = docs[n // 2 + 1]
    rows = torch.from_numpy(docs).to(DEV).to(TORCH_DT[storage])
    shard = ops.FlatShard(rows, first_global_id=1000)
    s, i = shard.search(torch.from_numpy(q).to(DEV), k, mode)
    s2, i2 = shard.search(torch.from_numpy(q).to(DEV), k, mode)          # cached segments and buffers: same answer
    s, i, s2, i2 = s.cpu().numpy(), i.cpu().numpy(), s2.cpu().numpy(), i2.cpu().numpy()
    assert np.array_equal(i, i2) and np.array_equal(s.view(np.int32), s2.view(np.int32))
    stored = rows.float().cpu().numpy()
    ws, wi = oracle.search(stored, q, k, oracle.CANONICAL, storage, 1000)
    if mode == "verify":
        assert np.array_equal(i, wi), np.argwhere(i != wi)[:4]
        assert np.array_equal(s.view(np.int32), ws.view(np.int32))
    else:
        assert recall(i, wi) >= 0.999
        fin = np.isfinite(ws)
        assert np.array_equal(np.isfinite(s), fin) and np.abs(np.sort(s[fin]) - np.sort(ws[fin])).max() <= 5e-3
    assert np.all(i[:, min(k, n):] == -1) and np.all(np.diff(s[:, :min(k, n)], axis=1) <= 0)


def test_merge_segments_c_abi_argument_errors():
    z = torch.zeros((65, 2, 128), device=DEV)
    zi = torch.zeros((65, 2, 128), dtype=torch.int64, device=DEV)
    with pytest.raises(ValueError):
        ops.merge_segments(z, zi, 200)                # 65 * 128 candidates > 8192
    with pytest.raises(ValueError):
        ops.merge_segments(z[:4].contiguous(), zi[:4].contiguous(), 1025)
    with pytest.raises(ValueError):
        ops.merge_segments(z[:4].contiguous(), zi[:4].contiguous().int(), 100)


def test_host_buffer_loop_two_stream_form_on_one_gpu():
    """ShardedFlat.search_host_pipelined(two_stream=True) without a process group: the H2D copy on a copy stream, the
    scan on the current stream, reduce + D2H on the side stream -- same answers as the synchronous host call, with
    three slots in flight."""
    from vietnamese_qa_system_b200.sharded import ShardedFlat
    rng = np.random.default_rng(9)
    docs = unit_rows(rng, 40000, 768)
    rows = torch.from_numpy(docs).to(DEV).to(torch.bfloat16)
    sh = ShardedFlat(rows, 40000, mode="fast")
    qs = [torch.from_numpy(unit_rows(rng, 32, 768)).pin_memory() for _ in range(4)]
    want = [tuple(t.clone() for t in sh.shard.search_host(q, 10, "fast")) for q in qs]
    pend = [None] * 3
    for step in range(11):
        slot = step % 3
        if pend[slot] is not None:
            hs, hi, ev, j = pend[slot]
            ev.synchronize()
            assert torch.equal(hi, want[j][1]) and torch.equal(hs, want[j][0])
        pend[slot] = sh.search_host_pipelined(qs[step % 4], 10, slot, two_stream=True) + (step % 4,)
    for hs, hi, ev, j in pend:
        ev.synchronize()
        assert torch.equal(hi, want[j][1]) and torch.equal(hs, want[j][0])
