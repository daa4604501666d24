#!/bin/bash
# Round-2 GPU call 3 (one B200): batched warm-up merges + dynamic tile schedule in the headline kernel, the register
# lists of the TS kernel out of local memory; what paces the TS kernel; the remaining test failures with tracebacks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
group() { echo "== $1"; shift; env "$@" CHECK=1 timeout 600 python tools/tune_worker.py 2>&1 | grep -v "^{" | tail -n 14; }
echo "== GPU tests (search + variants) with tracebacks"
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_experimental.py -m gpu -q --tb=short 2>&1 | tail -n 60
echo "== timeline of mma_topk_kernel at the 8-GPU shard size: dynamic vs static tiles"
ROWS=1250000 BATCHES=1,16,32 timeout 300 python tools/timeline_probe.py 2>&1 | tail -n 1 | tee $O/r2_timeline_shard_v2.json
ROWS=1250000 BATCHES=1,32 KNOBS=dyn_tiles=0 timeout 300 python tools/timeline_probe.py 2>&1 | tail -n 1 | tee $O/r2_timeline_shard_v2_static.json
group "headline kernel at the shard size" ROWS=1250000 K=10 MODE=tensor BATCHES=1,2,4,8,16,32 ITERS=50 "VARIANTS=-;VQA_DYN_TILES=0;-"
group "headline kernel at 10 M rows" ROWS=10000000 K=10 MODE=fast BATCHES=1,2,8,32 ITERS=10 "VARIANTS=-;VQA_DYN_TILES=0;VQA_STREAM_MAX_B=0"
group "large batches at the shard size" ROWS=1250000 K=10 MODE=fast BATCHES=64,128,256 ITERS=20 "VARIANTS=-;VQA_TS_KS=4"
echo "== what paces ts_topk_kernel"
ROWS=10000000 BATCHES=128,256 timeout 300 python tools/ts_waits_probe.py 2>&1 | tail -n 1 | tee $O/r2_ts_waits_10m.json
ROWS=1250000 BATCHES=64,128,256 timeout 300 python tools/ts_waits_probe.py 2>&1 | tail -n 1 | tee $O/r2_ts_waits_shard.json
ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 BATCHES=64 timeout 300 python tools/ts_waits_probe.py 2>&1 | tail -n 1 | tee $O/r2_ts_waits_cfgd.json
echo "== done"
