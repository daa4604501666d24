#!/bin/bash
# Round-2 GPU call 31 (one B200): where a launch of the 128-document-tile kernel goes at the 8-GPU shard size
# (1.25 M rows, B = 64 / 128 / 256): per-CTA life time against the role loops.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ROWS=1250000 BATCHES=64,128,256 MODE=fast timeout 300 python tools/ts_waits_probe.py 2>&1 | tail -n 1 | tee gpurun_out/r2_pair_waits_shard.json
echo "== done"
