#!/bin/bash
# large-batch sweep: side-by-side query chunks (VQA_MMA_GROUPS) at 10M and 1.25M rows
cd "$(dirname "$0")/.."
for ROWS in 10000000 1250000; do
  for G in 1 2 4 8; do
    echo -n "rows=$ROWS groups=$G -> "
    env ROWS=$ROWS MODE=tensor BATCHES=32,64,128,256 ITERS=5 VQA_MMA_GROUPS=$G python tools/tune_worker.py 2>&1 | grep GBps | tail -1
  done
done
