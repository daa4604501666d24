"""The search planner on the CPU (vqa_plan_describe: host arithmetic only, no device): which kernel family a
FAST search takes, and that EVERY planned launch fits the B200's shared memory (227 KB opt-in per CTA) and tensor
memory (512 columns) -- with the default routing and with the opt-in kernels (VQA_TS_QS, VQA_REDUCE_SELECT)
switched on.  A plan that does not fit would only fail on the GPU, as a launch error."""
import ctypes
import itertools

import pytest

from vietnamese_qa_system_b200 import _native as N

F32, BF16, F16 = 0, 1, 2
VERIFY, FAST, STREAM, TENSOR, TS = 0, 1, 2, 3, 4
SMS, SMEM = 148, 232448


def plan(n, dim, dtype, b, k, mode=FAST, **knobs):
    out = (ctypes.c_int32 * 16)()
    smem = ctypes.c_size_t()
    if knobs:
        t = N.tuning_default().update(**knobs)
        rc = N.lib().vqa_plan_describe_tuned(n, dim, dtype, b, k, mode, SMS, SMEM, ctypes.byref(t), out, ctypes.byref(smem))
    else:
        rc = N.lib().vqa_plan_describe(n, dim, dtype, b, k, mode, SMS, SMEM, out, ctypes.byref(smem))
    if rc != 0:
        return None
    keys = ["family", "pass_nq", "passes", "groups", "stages", "kps", "ncol", "split", "qs", "ks", "kscan", "k_out",
            "rescore", "tmem_query_cols"]
    d = dict(zip(keys, list(out)))
    d["smem"] = smem.value
    return d


KNOB_ENV = ("VQA_TS_QS", "VQA_TS_KS", "VQA_REDUCE_SELECT", "VQA_TS_SPLIT", "VQA_TS_EXTRA", "VQA_STREAM_MAX_B")


def test_default_routing(monkeypatch):
    """Round-2 defaults (flipped by the B200 timings in profiles/r2_call1.log): B <= 2 on the CUDA-core streaming
    kernel for shards of >= 8 GB, up to 32 queries on the smem-resident tcgen05 kernel, beyond that the QS variant of
    the TMEM-resident-query kernel with three accumulator stages at dim 768, and dim 1024 on it too."""
    for v in KNOB_ENV:
        monkeypatch.delenv(v, raising=False)
    n = 10_000_000
    for b in (1, 2):                                       # the CUDA-core streaming kernel is a knob away (it lost to the
        assert plan(n, 768, BF16, b, 10, stream_max_b=2)["family"] == STREAM         # seeded tcgen05 kernel at every size)
        assert plan(1_250_000, 768, BF16, b, 10, stream_max_b=2)["family"] == TENSOR  # ... and never on small shards
    for b in (1, 2, 3, 8, 16):                             # headline kernel: smem-resident tcgen05, hi/lo columns
        p = plan(n, 768, BF16, b, 10)
        assert p["family"] == TENSOR and p["split"] == 1 and p["smem"] <= SMEM
    for b in (17, 32):                                     # ... and for more than 16 queries over a scan of >= 12 GB screen
        p = plan(n, 768, BF16, b, 10)                      # mode (half the tensor work, exact re-score of k + 6): ss_screen = -1
        assert (p["family"], p["split"], p["kscan"], p["rescore"]) == (TENSOR, 0, 16, 1) and p["smem"] <= SMEM
        for rows, knobs in ((n, {"ss_screen": 0}), (5_000_000, {}), (1_250_000, {})):
            p = plan(rows, 768, BF16, b, 10, **knobs)      # forced off / the 2- and 8-GPU shards: hi/lo columns
            assert (p["family"], p["split"], p["rescore"]) == (TENSOR, 1, 0)
    assert plan(1_250_000, 768, BF16, 32, 10, ss_screen=1)["split"] == 0
    for b in (129, 256, 1024):                             # > 128 queries: CTA pairs (cta_group::2), 256 queries per launch
        p = plan(n, 768, BF16, b, 10)
        assert (p["family"], p["pass_nq"], p["ks"], p["kscan"], p["k_out"], p["rescore"]) == (5, 256, 4, 16, 32, 1)
        assert p["tmem_query_cols"] == 256 and p["smem"] <= SMEM and p["stages"] >= 3   # + two 128-column accumulators
    assert plan(n, 768, BF16, 256, 10, pair=0)["family"] == TS
    assert plan(n, 768, BF16, 256, 27)["family"] == TS      # k + spare > 32: no register lists -> TS kernel
    for b in (33, 64, 128):                                # 33..128 queries: the same 128-document tiles on single CTAs
        p = plan(n, 768, BF16, b, 10)
        assert (p["family"], p["pass_nq"], p["ks"], p["kscan"], p["rescore"]) == (5, 128, 4, 16, 1) and p["smem"] <= SMEM
        assert plan(n, 1024, BF16, b, 10)["family"] == TS  # ... at dim <= 768 only
    for b in (33, 64, 128):                                # knob off: TMEM-resident-query kernel, screen + re-score of 32
        p = plan(n, 768, BF16, b, 10, wide=0)
        assert (p["family"], p["split"], p["qs"], p["ks"], p["kscan"], p["k_out"], p["rescore"]) == (TS, 0, 1, 2, 16, 32, 1)
        assert p["tmem_query_cols"] == 320 and p["smem"] <= SMEM       # 10 blocks in TMEM -> 3 accumulator stages
    p = plan(n, 768, BF16, 8, 100)                         # k > 32: TS with hi/lo rows and heaps, no re-scoring
    assert (p["family"], p["split"], p["qs"], p["kscan"], p["rescore"]) == (TS, 1, 1, 100, 0)
    assert plan(n, 768, BF16, 1, 100)["family"] == TS      # big k is never a streaming-kernel case
    p = plan(12_500_000, 1024, F16, 64, 100)               # BASELINE configs[3]: one pass, 10 blocks in TMEM + 6 in smem,
    assert (p["family"], p["qs"], p["ks"], p["split"], p["rescore"]) == (TS, 1, 6, 0, 1)     # M = 64 instructions
    assert p["pass_nq"] == 64 and p["smem"] <= SMEM
    assert plan(n, 1024, BF16, 64, 10, TS) is not None
    assert plan(n, 768, F32, 4, 10)["family"] == STREAM and plan(n, 776, BF16, 4, 10)["family"] == STREAM
    assert plan(n, 768, BF16, 4, 10, VERIFY)["family"] == STREAM


def test_round1_routing_is_still_reachable_through_the_knobs(monkeypatch):
    for v in KNOB_ENV:
        monkeypatch.delenv(v, raising=False)
    r1 = dict(ts_qs=0, reduce_select=0, stream_max_b=0, wide=0, pair=0)
    n = 10_000_000
    assert plan(n, 768, BF16, 1, 10, **r1)["family"] == TENSOR
    p = plan(n, 768, BF16, 128, 10, **r1)
    assert (p["family"], p["split"], p["qs"], p["kscan"], p["k_out"], p["rescore"]) == (TS, 0, 0, 16, 32, 1)
    assert p["tmem_query_cols"] == 384 and p["stages"] * p["kps"] == 24
    assert plan(12_500_000, 1024, F16, 64, 100, **r1)["family"] == TENSOR     # dim 1024 does not fit TMEM without QS
    assert plan(n, 1024, BF16, 64, 10, TS, **r1) is None
    monkeypatch.setenv("VQA_TS_QS", "0")                   # the environment spelling is read by vqa_plan_describe too
    monkeypatch.setenv("VQA_WIDE", "0")
    assert plan(n, 768, BF16, 128, 10)["qs"] == 0


@pytest.mark.parametrize("qs,select", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_every_plan_fits_shared_and_tensor_memory(monkeypatch, qs, select):
    monkeypatch.setenv("VQA_TS_QS", str(qs))
    monkeypatch.setenv("VQA_REDUCE_SELECT", str(select))
    seen = set()
    dims = [64, 128, 192, 256, 384, 512, 640, 768, 832, 896, 960, 1024, 1088, 2048]
    for dim, dtype, b, k in itertools.product(dims, (BF16, F16), (1, 7, 32, 33, 64, 65, 128, 129, 300, 1024),
                                              (1, 10, 16, 26, 27, 32, 33, 64, 100, 122, 123, 128)):
        for ks in ([None] if not qs else [None, 0, 2, 4, 8, 16]):
            if ks is None:
                monkeypatch.delenv("VQA_TS_KS", raising=False)
            else:
                monkeypatch.setenv("VQA_TS_KS", str(ks))
            p = plan(2_000_000, dim, dtype, b, k)
            assert p is not None, (dim, dtype, b, k, ks)
            assert p["smem"] <= SMEM, (dim, dtype, b, k, ks, p)
            seen.add((p["family"], p["qs"], p["split"], p["rescore"]))
            if p["family"] == TS:
                kb = dim // 64
                assert p["stages"] >= 2 and kb % p["kps"] == 0 and 0 <= p["ks"] <= kb
                assert p["tmem_query_cols"] + 64 <= 512                       # at least one accumulator stage
                assert qs or (p["qs"] == 0 and p["ks"] == 0 and dim <= 768)
                assert p["kscan"] <= 128 and p["k_out"] <= 128 and p["kscan"] >= k
                if p["rescore"] and p["k_out"] > 32:                          # only the radix select re-scores > 32
                    assert select and qs and dtype == F16
                if dim > 768:
                    assert p["tmem_query_cols"] <= 384                        # two accumulator stages
            if p["family"] == TENSOR:
                assert p["stages"] >= 2 and p["ncol"] in (16, 32, 64, 128)
            if p["family"] == 5:                                              # 128-document tiles (pairs / single CTAs)
                kb = dim // 64
                assert b > 32 and k + 6 <= 32 and dim <= 1024 and (b > 128 or dim <= 768)
                assert p["stages"] >= 2 and kb % p["kps"] == 0 and 0 <= p["ks"] <= kb
                assert p["tmem_query_cols"] + 2 * 128 <= 512                  # two 128-column accumulator stages
                assert (p["kscan"], p["k_out"], p["rescore"]) == (k + 6, 32, 1)
    assert (TENSOR, 0, 1, 0) in seen and (5, 1, 0, 1) in seen
    if qs:
        assert (TS, 1, 0, 1) in seen and (TS, 1, 1, 0) in seen
    else:
        assert (TS, 0, 1, 0) in seen      # (screen-mode TS plans need the QS variant once 128-document tiles take dim <= 768)


def test_config_d_plan(monkeypatch):
    """BASELINE configs[3] (12.5 M x 1024 fp16 per GPU, B = 64, top-100) on the default routing:
    one pass of the TMEM-resident-query kernel -- 10 query blocks in tensor memory, 6 in shared memory, heaps for
    64 rows, 106 candidates per list, the 128 best re-scored exactly by the radix-select reduce."""
    monkeypatch.setenv("VQA_TS_QS", "1")
    monkeypatch.setenv("VQA_REDUCE_SELECT", "1")
    monkeypatch.delenv("VQA_TS_KS", raising=False)
    p = plan(12_500_000, 1024, F16, 64, 100)
    assert (p["family"], p["qs"], p["ks"], p["split"], p["kscan"], p["k_out"], p["rescore"]) == (TS, 1, 6, 0, 106, 128, 1)
    assert p["passes"] == 1 and p["tmem_query_cols"] == 320 and p["stages"] >= 3 and p["smem"] <= SMEM
    q = plan(12_500_000, 1024, BF16, 64, 100)              # bf16 rows: 8-bit queries would need ~28 spare ranks -> hi/lo rows
    assert (q["family"], q["qs"], q["split"], q["rescore"]) == (TS, 1, 1, 0) and q["smem"] <= SMEM
