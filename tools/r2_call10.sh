#!/bin/bash
# Round-2 GPU call 10 (one B200): the M = 64 variant after the heap-initialisation fix (tests, config D, B = 33..64),
# then the whole suite, smoke and the N = 1 bench line on the same build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
group() { echo "== $1"; shift; env "$@" CHECK=1 timeout 600 python tools/tune_worker.py 2>&1 | grep -v "^{\"" | tail -n 14; }
echo "== M = 64 variant tests"
timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q --tb=short -k "m64" 2>&1 | tail -n 15
group "config D shard: 12.5 M x 1024 fp16, B = 64, top-100" ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 ITERS=5 \
  "VARIANTS=-;VQA_TS_M64=1;VQA_TS_M64=1,VQA_TS_KS=4;VQA_TS_M64=1,VQA_TS_KS=8;VQA_TS_M64=1,VQA_TS_KS=10;VQA_TS_M64=1,VQA_TS_KS=16"
group "dim 768, 10 M rows, B = 33 / 48 / 64: TS M = 128, 128-document tiles, TS M = 64" ROWS=10000000 K=10 MODE=fast BATCHES=33,48,64 ITERS=10 \
  "VARIANTS=VQA_WIDE=0;-;VQA_WIDE=0,VQA_TS_M64=1;VQA_WIDE=0,VQA_TS_M64=1,VQA_TS_KS=4"
group "the same at the shard size" ROWS=1250000 K=10 MODE=fast BATCHES=33,48,64 ITERS=30 "VARIANTS=VQA_WIDE=0;-;VQA_WIDE=0,VQA_TS_M64=1"
group "dim 1024 fp16 top-10, B = 64" ROWS=8000000 DIM=1024 DTYPE=fp16 K=10 MODE=fast BATCHES=48,64 ITERS=5 "VARIANTS=-;VQA_TS_M64=1;VQA_TS_M64=1,VQA_TS_KS=8"
echo "== what paces the M = 64 kernel on config D"
ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 BATCHES=64 KNOBS=ts_m64=1 timeout 300 python tools/ts_waits_probe.py 2>&1 | tail -n 1 | tee $O/r2_ts_waits_cfgd_m64.json
echo "== full GPU test suite"
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 25 | tee $O/r2_pytest_gpu_final.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3
echo "== done"
