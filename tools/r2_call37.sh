#!/bin/bash
# Round-2 GPU call 37 (one B200): screen mode of the headline kernel as the planner's choice for long scans with more
# than 16 queries: the new parity test, the plan / family tests, then the N = 1 bench line (the driver's command).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_experimental.py tests/test_gpu_search.py -m gpu -q --tb=short -k "screen_mode_of_the_smem or family_selection or set_tuning or round_1_routing or graph" 2>&1 | tail -n 6
timeout 1200 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_bench_n1.json') if l.startswith('{')][-1])
print({k: d.get(k) for k in ('metric', 'value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k', 'gpu_launches', 'lib')})
print('roofline', {k: d['roofline'].get(k) for k in ('frac', 'step_frac', 'kernel_ms', 'kernel')}, 'e2e', d['e2e'], d['clocks'])
print('independent', d['independent_check']); print('fast_vs_verify_max_rel_score_err', d.get('fast_vs_verify_max_rel_score_err'))
for r in d.get('sweep') or []: print(r.get('batch'), round(r.get('ms', 0), 4), round(r.get('scan_ms', 0), 4), round(r.get('hbm_frac', 0), 3), str(r.get('family'))[:60])
PY
echo "== done"
