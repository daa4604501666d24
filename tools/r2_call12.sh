#!/bin/bash
# Round-2 GPU call 12 (one B200): the two-stream search (scan stream + reduce stream): tests, then the pipelined loop
# against the one-stream loop at the 8-GPU shard size and at 10 M rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== two-stream tests"
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -q --tb=short -k "two_stream or graph or host" 2>&1 | tail -n 15
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k')}, 'roofline', {k: d['roofline'][k] for k in ('frac', 'step_frac', 'kernel_ms')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['one_at_a_time_ms_per_step'])
PY
}
for b in 1 32; do
  echo "== bench.py --rows 1250000 --batch $b (pipelined loop on one shard)"
  timeout 300 python bench.py --rows 1250000 --batch $b --steps 300 --warmup 10 --sweep 0 --check 0 --no-cpu > $O/r2_bench_shard_b$b.json 2> $O/r2_bench_shard.err; tail -c 300 $O/r2_bench_shard.err; show $O/r2_bench_shard_b$b.json
done
echo "== bench.py N = 1 (10 M rows)"
timeout 600 python bench.py --steps 50 --warmup 5 --sweep 0 --no-cpu > $O/r2_bench_n1_2s.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err; show $O/r2_bench_n1_2s.json
echo "== done"
