"""K1 (fused masked mean-pool + L2 normalise), row normalise, merge-top-k and the agreement
rule against the oracle.  Floating point: |err| <= 1e-5 relative to the row's largest entry."""
import numpy as np
import pytest
import torch

import oracle
from vietnamese_qa_system_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def close(a, b):
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-6)
    return np.all(np.abs(a - b) <= TOL * scale)


def test_kat6_pool_golden(golden):
    h, m = torch.from_numpy(golden["kat6_hidden"]).to(DEV), torch.from_numpy(golden["kat6_mask"]).to(DEV)
    out = ops.pool_normalize(h, m, normalize=False).cpu().numpy()
    assert close(out, golden["kat6_pooled"]) and np.all(out[2] == 0)
    out = ops.pool_normalize(h, m, normalize=True).cpu().numpy()
    assert close(out, golden["kat6_pooled_norm"]) and np.all(out[2] == 0)


def test_kat7_normalise_golden(golden):
    x = torch.from_numpy(golden["kat7_x"]).to(DEV)
    out = ops.normalize_rows(x).cpu().numpy()
    assert close(out, golden["kat7_norm"]) and np.all(out[3] == 0) and np.all(np.isfinite(out))


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("mdt", [torch.int64, torch.int32, torch.bool, torch.float32])
@pytest.mark.parametrize("b,s,d", [(4, 16, 768), (3, 37, 384), (2, 1, 1024), (1, 128, 768), (5, 300, 64), (2, 7, 4096)])
def test_pool_normalize_vs_oracle(dt, mdt, b, s, d):
    g = torch.Generator().manual_seed(b * 1000 + s)
    h = torch.randn(b, s, d, generator=g).to(dt)
    lens = torch.randint(0, s + 1, (b,), generator=g)
    lens[0] = s
    mask = (torch.arange(s)[None, :] < lens[:, None])
    for norm in (True, False):
        out = ops.pool_normalize(h.to(DEV), mask.to(mdt).to(DEV), normalize=norm).cpu().numpy()
        ref = oracle.mean_pool(h.float().numpy(), mask.float().numpy(), norm)
        assert close(out, ref)
        ref2 = oracle.np_mean_pool(h.float().numpy(), mask.float().numpy(), norm)
        assert close(out, ref2)


def test_pool_config_e_shape_properties():
    """BASELINE configs[4]: [256,256,768] bf16, valid lengths U[16,256]: unit-norm output,
    padded tokens have no influence."""
    g = torch.Generator().manual_seed(5)
    h = torch.randn(256, 256, 768, generator=g).to(torch.bfloat16)
    lens = torch.randint(16, 257, (256,), generator=g)
    mask = (torch.arange(256)[None, :] < lens[:, None]).to(torch.int64)
    out = ops.pool_normalize(h.to(DEV), mask.to(DEV))
    assert torch.allclose(out.norm(dim=1), torch.ones(256, device=DEV), atol=1e-5)
    h2 = h.clone()
    h2[mask == 0] = 123.0
    out2 = ops.pool_normalize(h2.to(DEV), mask.to(DEV))
    assert torch.equal(out, out2)
    ref = oracle.mean_pool(h[:8].float().numpy(), mask[:8].numpy(), True)
    assert close(out[:8].cpu().numpy(), ref)


@pytest.mark.parametrize("n,d", [(1, 4), (1000, 768), (333, 384), (17, 8192)])
def test_normalize_rows_and_cast(n, d):
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, d, generator=g) * 3
    ref = oracle.normalize_rows(x.numpy())
    out = ops.normalize_rows(x.to(DEV))
    assert close(out.cpu().numpy(), ref)
    for cdt in (torch.bfloat16, torch.float16):
        c = ops.normalize_rows(x.to(DEV), cast_dtype=cdt)
        assert c.dtype == cdt and torch.equal(c, out.to(cdt))     # the cast is round-to-nearest of the fp32 result


@pytest.mark.parametrize("lists,b,kin,kout", [(2, 3, 10, 10), (8, 32, 10, 10), (8, 64, 100, 100), (3, 1, 5, 12),
                                              (148, 4, 10, 10), (4, 2, 128, 128)])
def test_merge_topk_vs_oracle(lists, b, kin, kout):
    rng = np.random.default_rng(lists * 100 + b)
    cs = rng.standard_normal((lists, b, kin)).astype(np.float32)
    cs[:, :, ::3] = np.round(cs[:, :, ::3], 1)                      # plant score ties across lists
    ci = rng.permutation(lists * b * kin).reshape(lists, b, kin).astype(np.int64)
    ci[0, 0, -1] = -1
    cs[0, 0, -1] = -np.inf
    ms, mi = ops.merge_topk(torch.from_numpy(cs).to(DEV), torch.from_numpy(ci).to(DEV), kout)
    rs, ri = oracle.merge_topk(cs, ci, kout)
    assert np.array_equal(mi.cpu().numpy(), ri) and np.array_equal(ms.cpu().numpy(), rs)


def test_agree_rule_vs_oracle():
    rng = np.random.default_rng(3)
    n = 1000
    ia, ib = rng.integers(0, 5, n), rng.integers(0, 5, n)
    sa = rng.uniform(0, 0.5, n).astype(np.float32)
    sb = rng.uniform(0, 0.5, n).astype(np.float32)
    sa[:4], sb[:4] = [0.2, 0.25, 0.125, 0.25], [0.2, 0.2, 0.25, 0.125]
    ia[:4] = ib[:4] = 1
    acc, comb = ops.agree(torch.from_numpy(ia).to(DEV), torch.from_numpy(sa).to(DEV),
                          torch.from_numpy(ib).to(DEV), torch.from_numpy(sb).to(DEV), 0.4)
    want = [oracle.agree(int(a), float(x), int(b), float(y)) for a, x, b, y in zip(ia, sa, ib, sb)]
    assert acc.cpu().tolist() == want
    assert want[:4] == [True, True, False, False]
    np.testing.assert_allclose(comb.cpu().numpy(), sa.astype(np.float64) + sb.astype(np.float64), rtol=1e-6)
