#!/bin/bash
# Round-2 GPU call 5 (one B200): headline kernel after the pipelined slot read; pair-kernel ring geometry; reduce with
# four interleaved partial lists; whole GPU suite; bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
group() { echo "== $1"; shift; env "$@" CHECK=1 timeout 600 python tools/tune_worker.py 2>&1 | grep -v "^{\"" | tail -n 14; }
echo "== full GPU test suite"
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 30 | tee $O/r2_pytest_gpu_call5.log
echo "== timeline of mma_topk_kernel at the 8-GPU shard size"
ROWS=1250000 BATCHES=16,32 timeout 300 python tools/timeline_probe.py 2>&1 | tail -n 1 | tee $O/r2_timeline_shard_v4.json | cut -c1-3000
group "headline kernel at the shard size" ROWS=1250000 K=10 MODE=fast BATCHES=1,2,4,8,16,32 ITERS=50 "VARIANTS=-;VQA_REDUCE_EARLY=0;-"
group "large batches at the shard size (pair for > 128)" ROWS=1250000 K=10 MODE=fast BATCHES=64,128,256,512 ITERS=20 "VARIANTS=-;VQA_PAIR=0"
group "pair kernel ring geometry, 10 M rows" ROWS=10000000 K=10 MODE=pair BATCHES=256 ITERS=5 "VARIANTS=-;VQA_MMA_KPS=2;VQA_MMA_KPS=1;VQA_MMA_KPS=3;VQA_MMA_KPS=6"
group "BASELINE configs[1]: 1 M x 768, B = 1, 32, 1024" ROWS=1000000 K=10 MODE=fast BATCHES=1,32,1024 ITERS=20 "VARIANTS=-;VQA_PAIR=0"
echo "== launch list at the shard size, B = 32"
ROWS=1250000 K=10 MODE=fast BATCHES=32 ITERS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
  -c 40 --csv --log-file $O/r2_launches_shard_b32_v2.csv python tools/tune_worker.py > /dev/null 2>&1
tail -n 6 $O/r2_launches_shard_b32_v2.csv | cut -c1-60,150-260
echo "== bench.py N = 1"
timeout 900 python bench.py --steps 50 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 400 $O/r2_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','one_step_at_a_time_ms','recall_at_k')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['one_at_a_time_ms_per_step'])
print(d['independent_check'])
for r in d['sweep']: print(r['batch'], round(r['ms'],4), round(r['scan_ms'],4), round(r['hbm_frac'],3), round(r['tensor_frac'],3), r['family'][:30])
print(d['pool_k1'])
PY
echo "== done"
