#!/bin/bash
# Round-2 GPU call 7 (one B200): 128-document tiles on single CTAs (B = 33..128); sustained-load behaviour of the
# headline kernel at B = 32 (hi/lo columns vs screen mode); K1 profile of the shipped build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
group() { echo "== $1"; shift; env "$@" CHECK=1 timeout 600 python tools/tune_worker.py 2>&1 | grep -v "^{\"" | tail -n 14; }
echo "== pair / wide-tile tests"
timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q --tb=short -k "pair or set_tuning" 2>&1 | tail -n 25
group "B = 64, 128 at 10 M rows: TS kernel vs 128-document tiles on single CTAs" ROWS=10000000 K=10 MODE=fast BATCHES=64,128 ITERS=10 "VARIANTS=-;VQA_WIDE=1;-;VQA_WIDE=1"
group "B = 64, 128 at the shard size" ROWS=1250000 K=10 MODE=fast BATCHES=48,64,128 ITERS=30 "VARIANTS=-;VQA_WIDE=1"
group "dim 1024 fp16, B = 64, 128 (top-10)" ROWS=8000000 DIM=1024 DTYPE=fp16 K=10 MODE=fast BATCHES=64,128 ITERS=5 "VARIANTS=-;VQA_WIDE=1"
group "sustained load, 10 M rows (200 iterations each)" ROWS=10000000 K=10 MODE=fast BATCHES=16,32 ITERS=200 "VARIANTS=-;VQA_SS_SCREEN=1;-;VQA_SS_SCREEN=1"
group "screen mode at the shard size" ROWS=1250000 K=10 MODE=fast BATCHES=16,32 ITERS=100 "VARIANTS=-;VQA_SS_SCREEN=1"
echo "== clocks / power while the B = 32 search loops for ~6 s"
(nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active,temperature.gpu --format=csv,noheader -lms 500 > $O/r2_smi_b32.log &)
ROWS=10000000 K=10 MODE=fast BATCHES=32 ITERS=2500 timeout 120 python tools/tune_worker.py 2>&1 | tail -n 1
kill %1 2>/dev/null; sleep 0.5; awk 'NR%2==1' $O/r2_smi_b32.log | tail -n 16
echo "== ncu: K1 (fused mean-pool + normalise) of the shipped build"
timeout 300 ncu --set full --clock-control none -k "regex:pool_normalize" -s 4 -c 1 -f -o $O/r2_pool_k1 python - <<'PY' > $O/r2_pool_k1.log 2>&1
import torch, sys
sys.path.insert(0, '.')
from vietnamese_qa_system_b200 import ops
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(5)
hidden = torch.randn((256, 256, 768), generator=g, device=dev).to(torch.bfloat16)
lens = torch.randint(16, 257, (256,), generator=g, device=dev)
mask = (torch.arange(256, device=dev)[None, :] < lens[:, None]).to(torch.int64)
for _ in range(6):
    ops.pool_normalize(hidden, mask)
torch.cuda.synchronize()
PY
python tools/ncu_summary.py $O/r2_pool_k1.ncu-rep > $O/r2_pool_k1.txt 2>&1; head -n 24 $O/r2_pool_k1.txt
echo "== done"
