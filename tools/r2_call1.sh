#!/bin/bash
# Round-2 GPU call 1 (one B200): the un-gated variant tests, every knob timed on the index shapes BASELINE names,
# the cta_group::2 hardware probe, and ncu captures of the kernels whose limiter is not yet shown by a counter.
#   gpurun --timeout 2100 -- 'bash tools/r2_call1.sh > gpurun_out/r2_call1.log 2>&1'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
group() {  # title, then ROWS=.. etc. and VARIANTS=..
  echo "== $1"; shift
  env "$@" CHECK=1 timeout 600 python tools/tune_worker.py 2>&1 | grep -v "^{" | tail -n 14
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader

echo "== full GPU test suite (variant tests included)"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 8 | tee $O/r2_pytest_gpu_call1.log

echo "== hardware probe: semantics of tcgen05.mma.cta_group::2"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o $O/cta2_probe tools/cta2_probe.cu 2>&1 | grep -i error
timeout 60 $O/cta2_probe 2>&1 | tail -n 14
timeout 60 $O/cta2_probe --alloc-leader-only 2>&1 | tail -n 14

group "headline kernel at the 8-GPU shard size (1.25 M rows): tournament bound (since removed), early-exit reduce" \
  ROWS=1250000 K=10 MODE=tensor BATCHES=1,2,8,16,32 ITERS=50 \
  "VARIANTS=-;VQA_MMA_TB=1;VQA_REDUCE_EARLY=1;VQA_MMA_TB=1,VQA_REDUCE_EARLY=1"
group "streaming kernel vs tcgen05 for B = 1, 2, 4 at the shard size" \
  ROWS=1250000 K=10 MODE=stream BATCHES=1,2,4 ITERS=50 "VARIANTS=-;VQA_REDUCE_EARLY=1"
group "large batches at the shard size: round-1 routing vs round-2 defaults vs ks" \
  ROWS=1250000 K=10 MODE=fast BATCHES=64,128,256,512 ITERS=20 \
  "VARIANTS=VQA_TS_QS=0,VQA_REDUCE_SELECT=0;VQA_TS_QS=0;-;VQA_TS_KS=2;VQA_TS_KS=6;VQA_TS_KS=0;VQA_PDL_CHAIN=1"
group "10 M x 768 bf16 top-10, B = 1..512 (FAST routing): round-1 routing vs round-2 defaults" \
  ROWS=10000000 K=10 MODE=fast BATCHES=1,2,32,64,128,256,512 ITERS=6 \
  "VARIANTS=VQA_TS_QS=0,VQA_REDUCE_SELECT=0,VQA_STREAM_MAX_B=0;-;VQA_TS_KS=2;VQA_TS_KS=6"
group "BASELINE configs[1]: 1 M x 768, B = 1, 32, 1024" \
  ROWS=1000000 K=10 MODE=fast BATCHES=1,32,1024 ITERS=20 \
  "VARIANTS=VQA_TS_QS=0,VQA_REDUCE_SELECT=0,VQA_STREAM_MAX_B=0;-;VQA_PDL_CHAIN=1"
group "top-100, 4 M x 768 bf16" \
  ROWS=4000000 K=100 MODE=fast BATCHES=8,64 ITERS=10 \
  "VARIANTS=VQA_TS_QS=0,VQA_REDUCE_SELECT=0;VQA_TS_QS=0;-;VQA_TS_SPLIT=0,VQA_TS_EXTRA=28"
group "BASELINE configs[3] shard: 12.5 M x 1024 fp16, B = 64, top-100 (HBM floor 3.9 ms)" \
  ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 ITERS=5 \
  "VARIANTS=-;VQA_TS_KS=6;VQA_TS_KS=8;VQA_TS_SPLIT=1;VQA_MMA_KPS=4;VQA_MMA_KPS=2;VQA_TS_EXTRA=12"
group "dim 1024 bf16 top-10, B = 32..256" \
  ROWS=8000000 DIM=1024 K=10 MODE=fast BATCHES=32,64,128,256 ITERS=5 "VARIANTS=VQA_TS_QS=0,VQA_REDUCE_SELECT=0;-;VQA_TS_KS=8"

echo "== ncu captures"
cap() {  # name, kernel regex, env...
  local name=$1 pat=$2; shift 2
  env "$@" ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s 3 -c 1 \
      -f -o $O/$name python tools/tune_worker.py > $O/$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/$name.txt 2>&1
  tail -n 34 $O/$name.txt
}
cap r2_mma_b32_shard mma_topk ROWS=1250000 K=10 MODE=fast BATCHES=32
cap r2_scan_b1_shard scan_topk ROWS=1250000 K=10 MODE=fast BATCHES=1
cap r2_ts_b256 ts_topk ROWS=10000000 K=10 MODE=fast BATCHES=256
cap r2_ts_b128 ts_topk ROWS=10000000 K=10 MODE=fast BATCHES=128
cap r2_ts_cfgd ts_topk ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64
cap r2_select_cfgd reduce_select ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64
echo "== launch list at the shard size, B = 32 (kernel shares of one search)"
ROWS=1250000 K=10 MODE=fast BATCHES=32 ITERS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
  -c 40 --csv --log-file $O/r2_launches_shard_b32.csv python tools/tune_worker.py > /dev/null 2>&1
tail -n 12 $O/r2_launches_shard_b32.csv | cut -c1-220
echo "== done"
