#!/bin/bash
# Round-2 GPU call 34 (one B200): the N = 1 bench line on the final tree (the driver's command: no flags).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_bench_n1.json') if l.startswith('{')][-1])
print({k: d.get(k) for k in ('metric', 'value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k', 'host_enqueue_us_per_step', 'gpu_launches', 'lib')})
print('roofline', {k: d['roofline'].get(k) for k in ('frac', 'step_frac', 'kernel_ms', 'kernel', 'traffic')}, 'e2e', d['e2e'], d['clocks'])
print('independent', d['independent_check']['ids_equal_independent'], 'cpu', d['cpu_baseline']['value'])
for r in d.get('sweep') or []: print(r.get('batch'), round(r.get('ms', 0), 4), round(r.get('hbm_frac', 0), 3), round(r.get('tensor_frac', 0), 3), str(r.get('family'))[:34])
print('pool_k1', d.get('pool_k1', {}).get('ms'), 'hybrid', (d.get('hybrid_leg') or {}).get('results_identical_to_cpu_restatement'))
PY
echo "== done"
