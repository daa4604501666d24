"""Small-shape run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).
usage: compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vietnamese_qa_system_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)


def unit(n, d):
    x = torch.randn(n, d, generator=g)
    return x / x.norm(dim=1, keepdim=True)


docs, q = unit(3000, 768), unit(300, 768)
for dt in (torch.float32, torch.bfloat16):
    shard = ops.FlatShard(docs.to(dev).to(dt))
    modes = ["verify", "stream"] + ([] if dt == torch.float32 else ["tensor", "ts", "pair", "fast"])
    for mode in modes:
        for b, k in ((1, 10), (9, 5), (32, 10), (40, 10), (70, 100), (130, 10), (300, 10)):
            if mode in ("verify", "stream") and b > 9:
                continue
            if mode == "pair" and k > 26:
                continue
            s, i = shard.search(q[:b].to(dev), k, mode)
            torch.cuda.synchronize()
            assert int((i >= 0).sum()) == b * min(k, 3000), (mode, b, k)
shard = ops.FlatShard(docs.to(dev).to(torch.bfloat16))
for mode in ("verify", "fast"):                                  # k > 128: segment searches + vqa_merge_segments
    s, i = shard.search(q[:3].to(dev), 700, mode)
    torch.cuda.synchronize()
    assert int((i >= 0).sum()) == 3 * 700, mode
h = torch.randn(5, 33, 768, generator=g).to(torch.bfloat16).to(dev)
m = (torch.arange(33)[None, :] < torch.tensor([33, 1, 0, 17, 8])[:, None]).to(torch.int64).to(dev)
out = ops.pool_normalize(h, m)
x = ops.normalize_rows(torch.randn(100, 768, generator=g).to(dev), cast_dtype=torch.bfloat16)
cs = torch.randn(8, 6, 10, generator=g).to(dev)
ci = torch.randint(0, 10 ** 6, (8, 6, 10), generator=g).to(dev)
ops.merge_topk(cs, ci, 10)
ops.agree(ci[0, 0], cs[0, 0], ci[1, 0], cs[1, 0])
torch.cuda.synchronize()
print("sanitize run complete")
