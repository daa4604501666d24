#!/bin/bash
# Round-2 GPU call 4 (one B200): warm-up seed + dynamic tiles v2 in the headline kernel; first run of the CTA-pair
# (cta_group::2) kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
group() { echo "== $1"; shift; env "$@" CHECK=1 timeout 600 python tools/tune_worker.py 2>&1 | grep -v "^{" | tail -n 14; }
echo "== GPU tests (search + variants + pair) with tracebacks"
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_experimental.py -m gpu -q --tb=short 2>&1 | tail -n 40
echo "== timeline of mma_topk_kernel at the 8-GPU shard size"
ROWS=1250000 BATCHES=1,16,32 timeout 300 python tools/timeline_probe.py 2>&1 | tail -n 1 | tee $O/r2_timeline_shard_v3.json | cut -c1-5000
group "headline kernel at the shard size" ROWS=1250000 K=10 MODE=tensor BATCHES=1,2,4,8,16,32 ITERS=50 \
  "VARIANTS=-;VQA_SEED=0;VQA_DYN_TILES=0;VQA_SEED=0,VQA_DYN_TILES=0;-"
group "headline kernel at 10 M rows" ROWS=10000000 K=10 MODE=tensor BATCHES=1,8,32 ITERS=10 "VARIANTS=-;VQA_SEED=0,VQA_DYN_TILES=0;-"
group "CTA-pair kernel vs TS kernel, 10 M rows" ROWS=10000000 K=10 MODE=pair BATCHES=256,512 ITERS=5 "VARIANTS=-;VQA_TS_KS=6;VQA_TS_KS=8"
group "TS kernel, 10 M rows" ROWS=10000000 K=10 MODE=ts BATCHES=256,512 ITERS=5 "VARIANTS=-"
group "CTA-pair kernel, 1.25 M-row shard" ROWS=1250000 K=10 MODE=pair BATCHES=256,512 ITERS=20 "VARIANTS=-"
group "CTA-pair kernel, 1 M rows (BASELINE configs[1])" ROWS=1000000 K=10 MODE=pair BATCHES=1024 ITERS=20 "VARIANTS=-"
echo "== what paces the pair kernel"
ROWS=10000000 BATCHES=256 MODE=pair timeout 300 python tools/ts_waits_probe.py 2>&1 | tail -n 1 | tee $O/r2_pair_waits_10m.json
echo "== ncu: pair kernel at B = 256"
ROWS=10000000 K=10 MODE=pair BATCHES=256 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:ts_pair" -s 3 -c 1 \
    -f -o $O/r2_pair_b256 python tools/tune_worker.py > $O/r2_pair_b256.log 2>&1
python tools/ncu_summary.py $O/r2_pair_b256.ncu-rep > $O/r2_pair_b256.txt 2>&1; head -n 24 $O/r2_pair_b256.txt
echo "== done"
