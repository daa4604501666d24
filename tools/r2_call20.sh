#!/bin/bash
# Round-2 GPU call 20 (one B200): config D's shard (12.5 M x 1024 fp16, B = 64, top-100): how many of the 16 query
# column blocks stay in tensor memory (ts_ks = blocks in shared memory; fewer TMEM columns = more accumulator stages,
# and SS-form instructions for those blocks) -- and the same for the ring width and extra ranks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 BATCHES=64,32 MODE=fast ITERS=15 CHECK=1 \
VARIANTS="-;ts_ks=4;ts_ks=8;ts_ks=10;ts_ks=12;ts_ks=14;ts_ks=16;ts_ks=8,mma_kps=2;ts_ks=10,mma_kps=2;ts_ks=12,mma_kps=2;ts_ks=12,mma_kps=1;ts_m64=0" \
timeout 900 python tools/tune_worker.py 2>&1 | tail -n 40
echo "== done"
