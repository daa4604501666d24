#!/bin/bash
# Round-2 GPU call 9 (one B200): final build -- whole GPU suite, smoke, sanitizers over every kernel family (incl. the
# CTA-pair kernel, warm-up seed, dynamic tiles), ncu at the north-star operating point, the bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== hardware probe: tensor-memory lanes of M = 64 instructions"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o $O/m64_probe tools/m64_probe.cu 2>&1 | grep -i error
timeout 60 $O/m64_probe 2>&1 | tail -n 40
echo "== M = 64 variant: tests, then config D's shard and B = 48 / 64 at dim 768"
timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q --tb=short -k "m64" 2>&1 | tail -n 25
group() { echo "== $1"; shift; env "$@" CHECK=1 timeout 600 python tools/tune_worker.py 2>&1 | grep -v "^{\"" | tail -n 14; }
group "config D shard: 12.5 M x 1024 fp16, B = 64, top-100" ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64 ITERS=5 \
  "VARIANTS=-;VQA_TS_M64=1;VQA_TS_M64=1,VQA_TS_KS=4;VQA_TS_M64=1,VQA_TS_KS=8;VQA_TS_M64=1,VQA_TS_KS=10"
group "dim 768, B = 48 / 64 (TS mode): M = 128 vs M = 64" ROWS=10000000 K=10 MODE=ts BATCHES=48,64 ITERS=10 "VARIANTS=-;VQA_TS_M64=1;VQA_TS_M64=1,VQA_TS_KS=0"
group "top-100 at dim 768, B = 8 / 64 (4 M rows)" ROWS=4000000 K=100 MODE=fast BATCHES=8,64 ITERS=10 "VARIANTS=-;VQA_TS_M64=1"
echo "== full GPU test suite"
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 25 | tee $O/r2_pytest_gpu_final.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3
echo "== compute-sanitizer"
for tool in memcheck racecheck synccheck; do
  echo "-- $tool"; timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run complete|Error|hazard" | head -n 12
done 2>&1 | tee $O/r2_sanitizer.txt
cap() {  # name, kernel regex, env...
  local name=$1 pat=$2; shift 2
  env "$@" ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s 3 -c 1 \
      -f -o $O/$name python tools/tune_worker.py > $O/$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/$name.txt 2>&1; head -n 9 $O/$name.txt
}
echo "== ncu at the north-star operating point and for the 128-document tiles"
cap r2_mma_b32_shard_final mma_topk ROWS=1250000 K=10 MODE=fast BATCHES=32
cap r2_mma_b1_shard_final mma_topk ROWS=1250000 K=10 MODE=fast BATCHES=1
cap r2_wide_b128_10m ts_pair ROWS=10000000 K=10 MODE=fast BATCHES=128
echo "== launch list of bench.py --steps 20 (kernel shares of the step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r2_launches_bench_n1.csv \
  python bench.py --steps 20 --warmup 3 --sweep 0 --check 0 --no-cpu > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2_launches_bench_n1.csv')))
hdr = next(r for r in rows if r and r[0] == 'ID')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if len(r) == len(hdr) and r[0] != 'ID':
        d = dict(zip(hdr, r)); name = d['Kernel Name'].split('(')[0][-60:]
        if 'vqa' in d['Kernel Name']:
            agg[name][0] += 1; agg[name][1] += float(d['Metric Value'])
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]: print(f"{k:62s} launches {n:5d}  total {t/1e6:9.3f} ms  avg {t/n/1e3:9.2f} us")
PY
echo "== bench.py N = 1"
timeout 900 python bench.py --steps 50 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_bench_n1.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k')}, 'roofline', {k: d['roofline'][k] for k in ('frac', 'step_frac', 'kernel_ms', 'traffic')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['one_at_a_time_ms_per_step'], d['clocks'])
print(d['independent_check'])
for r in d['sweep']: print(r['batch'], round(r['ms'], 4), round(r['scan_ms'], 4), round(r['hbm_frac'], 3), round(r['tensor_frac'], 3), r['family'][:60])
print(d['pool_k1']); print(d['config_a_reference_scale']); print(d['hybrid_leg'])
PY
echo "== bench.py --config D (one shard)"
timeout 600 python bench.py --config D --steps 10 --warmup 3 --no-cpu > $O/r2_bench_cfgd_n1.json 2> $O/r2_bench_cfgd.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_bench_cfgd_n1.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k')}, 'roofline', {k: d['roofline'][k] for k in ('frac', 'step_frac', 'kernel_ms', 'traffic')}, d['independent_check']['verify_vs_independent'])
PY
echo "== done"
