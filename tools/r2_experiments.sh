#!/bin/bash
# Round-2 first GPU call: time the opt-in paths that so far only ran on the CPU emulator, next to the defaults.
#   gpurun --timeout 1800 -- 'bash tools/r2_experiments.sh > gpurun_out/r2_experiments.log 2>&1'
# One process per index shape; inside it every knob setting ("variant", "-" = defaults) is timed on the same
# generated index:  variant -> {batch: {ms, GBps, recall, max_rel_err}}  (CUDA events around vqa_search; recall and
# score error against the fp32 verify kernel).  Roughly 20 GPU-minutes in all.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
group() {  # title, then ROWS=.. etc. and VARIANTS=..
  echo "== $1"; shift
  env "$@" CHECK=1 timeout 900 python tools/tune_worker.py 2>&1 | grep -v "^{" | tail -n 12
}

echo "== correctness first: the experimental GPU tests"
VQA_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_experimental.py -x -q 2>&1 | tail -n 5

group "headline kernel at the 8-GPU shard size (1.25 M rows): tournament bound, early-exit reduce (B-dependent cost)" \
  ROWS=1250000 K=10 MODE=tensor BATCHES=1,8,16,32 ITERS=50 \
  "VARIANTS=-;VQA_MMA_TB=1;VQA_REDUCE_EARLY=1;VQA_MMA_TB=1,VQA_REDUCE_EARLY=1"

group "headline kernel, full size (10 M rows)" \
  ROWS=10000000 K=10 MODE=tensor BATCHES=1,32 ITERS=10 "VARIANTS=-;VQA_MMA_TB=1,VQA_REDUCE_EARLY=1"

group "top-100, 4 M x 768 bf16: radix select; QS hi/lo heaps with early accumulator release; screen with 128 candidates" \
  ROWS=4000000 K=100 MODE=fast BATCHES=8,64 ITERS=10 \
  "VARIANTS=-;VQA_REDUCE_SELECT=1;VQA_REDUCE_SELECT=1,VQA_TS_QS=1;VQA_REDUCE_SELECT=1,VQA_TS_QS=1,VQA_TS_SPLIT=0,VQA_TS_EXTRA=28"

group "BASELINE configs[3] shard: 12.5 M x 1024 fp16, B = 64, top-100 (HBM floor 3.9 ms)" \
  ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 ITERS=5 \
  "VARIANTS=-;VQA_REDUCE_SELECT=1;VQA_REDUCE_SELECT=1,VQA_TS_QS=1;VQA_REDUCE_SELECT=1,VQA_TS_QS=1,VQA_TS_KS=6;VQA_REDUCE_SELECT=1,VQA_TS_QS=1,VQA_TS_KS=8;VQA_REDUCE_SELECT=1,VQA_TS_QS=1,VQA_TS_SPLIT=1"

group "dim 1024 bf16 top-10, B = 32..256 (default: smem-resident kernel with multicast groups)" \
  ROWS=8000000 DIM=1024 K=10 MODE=fast BATCHES=32,64,128,256 ITERS=5 "VARIANTS=-;VQA_TS_QS=1;VQA_TS_QS=1,VQA_REDUCE_SELECT=1"

group "10 M x 768 bf16 top-10, B = 64..512: re-scoring reduce (select kernel), accumulator stages (ks 0/2/4/6 -> 2/3/4/5)" \
  ROWS=10000000 K=10 MODE=fast BATCHES=64,128,256,512 ITERS=5 \
  "VARIANTS=-;VQA_REDUCE_SELECT=1;VQA_TS_QS=1,VQA_TS_KS=0;VQA_TS_QS=1,VQA_TS_KS=2;VQA_TS_QS=1,VQA_TS_KS=4;VQA_TS_QS=1,VQA_TS_KS=6;VQA_TS_QS=1,VQA_TS_KS=4,VQA_REDUCE_SELECT=1"

group "the same large batches at the 8-GPU shard size (fixed costs dominate)" \
  ROWS=1250000 K=10 MODE=fast BATCHES=64,128,256 ITERS=20 \
  "VARIANTS=-;VQA_REDUCE_SELECT=1;VQA_TS_QS=1,VQA_TS_KS=4,VQA_REDUCE_SELECT=1"

group "BASELINE configs[1]: 1 M x 768, B = 512 / 1024 (2 / 4 launches): scan i+1 overlapping reduce i" \
  ROWS=1000000 K=10 MODE=fast BATCHES=512,1024 ITERS=20 \
  "VARIANTS=-;VQA_PDL_CHAIN=1;VQA_PDL_CHAIN=1,VQA_REDUCE_SELECT=1;VQA_PDL_CHAIN=1,VQA_REDUCE_SELECT=1,VQA_TS_QS=1,VQA_TS_KS=4"

echo "== end-to-end with host buffers: synchronous calls vs two batches in flight"
ROWS=10000000 BATCH=32 timeout 600 python tools/e2e_pipeline_probe.py 2>&1 | tail -n 1
ROWS=1250000 BATCH=32 STEPS=1000 timeout 600 python tools/e2e_pipeline_probe.py 2>&1 | tail -n 1

echo "== hardware probe: semantics of tcgen05.mma.cta_group::2 (next step for the tensor-bound regime, DESIGN 8.3)"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o gpurun_out/cta2_probe tools/cta2_probe.cu 2>&1 | grep -i error
timeout 60 gpurun_out/cta2_probe 2>&1 | tail -n 12
timeout 60 gpurun_out/cta2_probe --alloc-leader-only 2>&1 | tail -n 12

echo "== regression check of the default path: the headline bench line"
timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -n 1 | cut -c1-600
