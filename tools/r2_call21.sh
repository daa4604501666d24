#!/bin/bash
# Round-2 GPU call 21 (one B200): the record of the final build -- whole GPU suite, smoke, sanitizers, ncu of the
# headline kernel (10 M rows, B = 32) and of config D's scan, launch list of the bench, the N = 1 bench line, config D
# on one shard, the CPU reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== full GPU test suite"
timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 25 | tee $O/r2_pytest_gpu_final.log
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
echo "== ncu of the final build: headline kernel at 10 M rows, config D's scan"
cap r2_mma_b32_10m_final mma_topk ROWS=10000000 K=10 MODE=fast BATCHES=32
cap r2_ts_cfgd_final ts_topk ROWS=12500000 DIM=1024 DTYPE=fp16 K=100 MODE=fast BATCHES=64
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
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: d.get(k) for k in ('metric', 'value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k', 'host_enqueue_us_per_step')})
print('roofline', {k: d['roofline'].get(k) for k in ('frac', 'step_frac', 'kernel_ms', 'kernel', 'traffic')}, 'e2e', d['e2e'], d['clocks'])
print('independent', d.get('independent_check')); print('cpu', d.get('cpu_baseline'))
for r in d.get('sweep') or []: print(r.get('batch'), round(r.get('ms', 0), 4), round(r.get('scan_ms', 0), 4), round(r.get('hbm_frac', 0), 3), round(r.get('tensor_frac', 0), 3), str(r.get('family'))[:34])
print('pool_k1', d.get('pool_k1'))
PY
}
echo "== bench.py N = 1 (the driver's command: no flags)"
timeout 1200 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err; show $O/r2_bench_n1.json
echo "== bench.py --config D on one shard (12.5 M x 1024 fp16, B = 64, top-100)"
timeout 900 python bench.py --config D --steps 20 --warmup 3 --sweep 0 > $O/r2_bench_cfgd_n1.json 2> $O/r2_bench_cfgd_n1.err; tail -c 300 $O/r2_bench_cfgd_n1.err; show $O/r2_bench_cfgd_n1.json
echo "== bench.py --impl reference (CPU arm, full 10 M-row steps)"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err; tail -c 300 $O/r2_bench_reference_arm.err; cat $O/r2_bench_reference_arm.json | cut -c1-900
echo "== done"
