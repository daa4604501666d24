#!/bin/bash
# Round-2 second GPU call: ncu --set full captures (one launch each, --import-source on) of the kernels whose
# limiter is not yet explained by a counter, summarised with tools/ncu_summary.py into gpurun_out/r2_*.txt.
#   gpurun --timeout 1500 -- 'bash tools/r2_profile.sh > gpurun_out/r2_profile.log 2>&1'
# Read here with:  ncu -i gpurun_out/<name>.ncu-rep --page source --csv   (per-line stall reasons; -lineinfo is on)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name, kernel regex, env...
  local name=$1 pat=$2; shift 2
  env "$@" ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s 3 -c 1 \
      -f -o gpurun_out/$name python tools/tune_worker.py > gpurun_out/$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1
  tail -n 40 gpurun_out/$name.txt
}
# 1. tensor-bound regime: B = 256 (cluster of 2, multicast) -- tensor pipe 58 % of sustained peak in round 1: why?
cap r2_ts_b256 ts_topk ROWS=10000000 K=10 MODE=fast BATCHES=256
cap r2_ts_b256_qs4 ts_topk ROWS=10000000 K=10 MODE=fast BATCHES=256 VQA_TS_QS=1 VQA_TS_KS=4
# 2. one chunk of 128 without a cluster: isolates the pair lock-step from the epilogue
cap r2_ts_b128_qs4 ts_topk ROWS=10000000 K=10 MODE=fast BATCHES=128 VQA_TS_QS=1 VQA_TS_KS=4
# 3. the headline kernel at the 8-GPU shard size (1.25 M rows, B = 32): prologue / threshold warm-up share
cap r2_mma_b32_shard mma_topk ROWS=1250000 K=10 MODE=fast BATCHES=32
# 4. config D shard on the QS path and the radix-select reduce behind it
cap r2_ts_cfgd ts_topk ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 VQA_TS_QS=1 VQA_REDUCE_SELECT=1
cap r2_select_cfgd reduce_select ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 VQA_TS_QS=1 VQA_REDUCE_SELECT=1
