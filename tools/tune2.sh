#!/bin/bash
# second tuning pass: kernel family x batch at the 1-GPU (10M) and 8-GPU-shard (1.25M) sizes
cd "$(dirname "$0")/.."
for ROWS in 10000000 1250000; do
  for cfg in "MODE=tensor BATCHES=1,4,8,16,32" "MODE=tensor BATCHES=8,32 VQA_MMA_KPS=1" "MODE=tensor BATCHES=8,32 VQA_MMA_KPS=3" "MODE=tensor BATCHES=8,16 VQA_MMA_KPS=6" "MODE=stream BATCHES=1,2,4,8"; do
    echo -n "rows=$ROWS $cfg -> "
    env ROWS=$ROWS $cfg python tools/tune_worker.py 2>&1 | grep GBps | tail -1
  done
done
