"""Full-size properties (sizes the CPU oracle cannot replay in seconds): 1M x 768 bf16 on one GPU."""
import numpy as np
import pytest
import torch

from vietnamese_qa_system_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N, D = 1_000_000, 768


@pytest.fixture(scope="module")
def big():
    g = torch.Generator(device=DEV).manual_seed(1234)
    rows = torch.empty((N, D), dtype=torch.bfloat16, device=DEV)
    for lo in range(0, N, 250_000):
        rows[lo:lo + 250_000] = ops.normalize_rows(torch.randn((250_000, D), generator=g, device=DEV),
                                                   cast_dtype=torch.bfloat16)
    # planted exact duplicates (seed 7) to exercise the tie rule at scale
    gp = torch.Generator(device="cpu").manual_seed(7)
    pairs = torch.randint(0, N, (1000, 2), generator=gp)
    rows[pairs[:, 1].to(DEV)] = rows[pairs[:, 0].to(DEV)]
    q = ops.normalize_rows(torch.randn((64, D), generator=g, device=DEV))
    q[:16] = rows[pairs[:16, 0].to(DEV)].float()       # queries that hit a duplicated pair exactly
    return rows, ops.normalize_rows(q), pairs


def test_kernel_families_agree_with_verify(big):
    rows, q, _ = big
    shard = ops.FlatShard(rows)
    sv, iv = shard.search(q, 10, "verify")
    for mode in ("stream", "tensor", "ts", "fast"):
        s, i = shard.search(q, 10, mode)
        rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(i.cpu(), iv.cpu())])
        assert rec >= 0.999, (mode, rec)
        assert torch.all((s - sv).abs() <= 1e-5 * sv.abs() + 2e-6)


def test_duplicate_pairs_come_back_lower_id_first(big):
    rows, q, pairs = big
    shard = ops.FlatShard(rows)
    for mode in ("verify", "tensor"):
        s, i = shard.search(q[:16], 2, mode)
        i = i.cpu()
        for r in range(16):
            a, b = sorted(pairs[r].tolist())
            if a != b:
                assert i[r].tolist() == [a, b], (mode, r)
        assert torch.all(s[:, 0] == s[:, 1])


def test_shard_concat_equals_merge_verify_bit_identical(big):
    """Result is bit-identical for 1/2/4/8-way row sharding (contiguous shards + id-asc tie rule)."""
    rows, q, _ = big
    full_s, full_i = ops.FlatShard(rows).search(q, 10, "verify")
    for g in (2, 4, 8):
        per = -(-N // g)
        cs, ci = [], []
        for r in range(g):
            sh = ops.FlatShard(rows[r * per:min((r + 1) * per, N)], first_global_id=r * per)
            s, i = sh.search(q, 10, "verify")
            cs.append(s)
            ci.append(i)
        ms, mi = ops.merge_topk(torch.stack(cs), torch.stack(ci), 10)
        assert torch.equal(mi, full_i) and torch.equal(ms, full_s)


def test_idempotent_and_batch_independent(big):
    rows, q, _ = big
    shard = ops.FlatShard(rows)
    s1, i1 = shard.search(q[:32], 10, "tensor")
    s2, i2 = shard.search(q[:32], 10, "tensor")
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    s3, i3 = shard.search(q[:8], 10, "tensor")          # a query's answer does not depend on its batch
    assert torch.equal(i3, i1[:8]) and torch.all((s3 - s1[:8]).abs() <= 2e-6)
    sv, iv = shard.search(q[:3], 10, "verify")
    sv1 = torch.cat([shard.search(q[j:j + 1], 10, "verify")[0] for j in range(3)])
    assert torch.equal(sv, sv1)


def test_top1_of_self_query_is_self(big):
    rows, _, _ = big
    picks = torch.tensor([0, 127, 128, 999_999, 500_000], device=DEV)
    q = ops.normalize_rows(rows[picks].float())
    s, i = ops.FlatShard(rows).search(q, 1, "fast")
    assert torch.all(s[:, 0] > 0.999)
    assert torch.all(rows[i[:, 0]].float().sub(rows[picks].float()).abs().max(dim=1).values == 0)
