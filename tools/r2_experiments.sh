#!/bin/bash
# Round-2 first GPU call: time the opt-in paths that so far only ran on the CPU emulator, next to the defaults.
#   gpurun --timeout 1500 -- 'bash tools/r2_experiments.sh > gpurun_out/r2_experiments.log 2>&1'
# Each line: knobs -> {batch: {ms, GBps, recall, max_rel_err}} (CUDA events around vqa_search, recall vs verify mode).
cd "$(dirname "$0")/.."
run() { echo -n "$* -> "; env "$@" CHECK=1 timeout 600 python tools/tune_worker.py 2>&1 | tail -n 1; }

echo "== correctness first: the experimental GPU tests"
VQA_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -x -q 2>&1 | tail -n 5

echo "== headline kernel, 8-GPU shard size (1.25 M rows) and full size: tournament bound off / on (B-dependent list-update cost)"
for TB in 0 1; do run ROWS=1250000 K=10 MODE=tensor BATCHES=1,8,16,32 ITERS=50 VQA_MMA_TB=$TB; done
for TB in 0 1; do run ROWS=10000000 K=10 MODE=tensor BATCHES=1,32 ITERS=10 VQA_MMA_TB=$TB; done

echo "== same, with the early-exit reduce (and both)"
run ROWS=1250000 K=10 MODE=tensor BATCHES=1,8,16,32 ITERS=50 VQA_REDUCE_EARLY=1
run ROWS=1250000 K=10 MODE=tensor BATCHES=1,8,16,32 ITERS=50 VQA_REDUCE_EARLY=1 VQA_MMA_TB=1

echo "== top-100, 4M x 768 bf16: list-insertion reduce vs radix select (default kernel family = TS hi/lo heaps)"
for SEL in 0 1; do run ROWS=4000000 K=100 MODE=fast BATCHES=8,64 ITERS=10 VQA_REDUCE_SELECT=$SEL; done

echo "== top-100, 4M x 768 bf16 on the QS kernel: hi/lo heaps with early accumulator release; screen (128 candidates) + exact re-score"
run ROWS=4000000 K=100 MODE=fast BATCHES=8,64 ITERS=10 VQA_REDUCE_SELECT=1 VQA_TS_QS=1
run ROWS=4000000 K=100 MODE=fast BATCHES=8,64 ITERS=10 VQA_REDUCE_SELECT=1 VQA_TS_QS=1 VQA_TS_SPLIT=0 VQA_TS_EXTRA=28

echo "== BASELINE configs[3] shard: 12.5M x 1024 fp16, B = 64, top-100 (HBM floor 3.9 ms)"
run ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 ITERS=5
run ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 ITERS=5 VQA_REDUCE_SELECT=1
for KS in 4 6 8; do
  run ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 ITERS=5 VQA_REDUCE_SELECT=1 VQA_TS_QS=1 VQA_TS_KS=$KS
done
run ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 ITERS=5 VQA_REDUCE_SELECT=1 VQA_TS_QS=1 VQA_TS_SPLIT=1

echo "== dim 1024 bf16 top-10, B = 32..256 (default: smem-resident kernel with multicast groups)"
run ROWS=8000000 DIM=1024 K=10 MODE=fast BATCHES=32,64,128,256 ITERS=5
run ROWS=8000000 DIM=1024 K=10 MODE=fast BATCHES=32,64,128,256 ITERS=5 VQA_TS_QS=1

echo "== large batches: warp-per-query re-scoring reduce (63 us per 128 queries) vs CTA-per-query select kernel; full size and 8-GPU shard"
for SEL in 0 1; do run ROWS=10000000 K=10 MODE=fast BATCHES=64,128,256 ITERS=5 VQA_REDUCE_SELECT=$SEL; done
for SEL in 0 1; do run ROWS=1250000 K=10 MODE=fast BATCHES=64,128,256 ITERS=20 VQA_REDUCE_SELECT=$SEL; done

echo "== BASELINE configs[1]: 1M x 768, B = 1024 (4 launches of 256): scan i+1 overlapping reduce i; + select kernel"
run ROWS=1000000 K=10 MODE=fast BATCHES=512,1024 ITERS=20
run ROWS=1000000 K=10 MODE=fast BATCHES=512,1024 ITERS=20 VQA_PDL_CHAIN=1
run ROWS=1000000 K=10 MODE=fast BATCHES=512,1024 ITERS=20 VQA_PDL_CHAIN=1 VQA_REDUCE_SELECT=1

echo "== 10M x 768 bf16 top-10, B = 64..512: accumulator stages (ks = 0: 2 stages, 2: 3, 4: 4, 6: 5)"
run ROWS=10000000 K=10 MODE=fast BATCHES=64,128,256,512 ITERS=5
for KS in 0 2 4 6; do
  run ROWS=10000000 K=10 MODE=fast BATCHES=64,128,256,512 ITERS=5 VQA_TS_QS=1 VQA_TS_KS=$KS
done

echo "== end-to-end with host buffers: synchronous calls vs two batches in flight"
ROWS=10000000 BATCH=32 timeout 600 python tools/e2e_pipeline_probe.py 2>&1 | tail -n 1
ROWS=1250000 BATCH=32 STEPS=1000 timeout 600 python tools/e2e_pipeline_probe.py 2>&1 | tail -n 1

echo "== hardware probe: semantics of tcgen05.mma.cta_group::2 (next step for the tensor-bound regime, DESIGN 8.3)"
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o gpurun_out/cta2_probe tools/cta2_probe.cu 2>&1 | grep -i error
timeout 60 gpurun_out/cta2_probe 2>&1 | tail -n 12
timeout 60 gpurun_out/cta2_probe --alloc-leader-only 2>&1 | tail -n 12

echo "== regression check of the default path: the headline bench line"
timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -n 1 | cut -c1-600
