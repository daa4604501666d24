#!/bin/bash
# Round-2 GPU call 36 (one B200): the host-buffer loop on ONE GPU through the two-stream form against the single C-ABI
# call (vqa_search_host_async): test, then both in one bench run at 10 M rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_search.py -m gpu -q --tb=short -k "two_stream_form or host" 2>&1 | tail -n 5
timeout 600 python bench.py --steps 200 --warmup 10 --sweep 0 --check 0 --no-cpu > gpurun_out/r2_bench_n1_e2e.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_bench_n1_e2e.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms')}, 'e2e', d['e2e'], d['clocks'])
PY
echo "== done"
