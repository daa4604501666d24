#!/bin/bash
# Round-2 GPU call 32 (one B200): per-slot scan streams in the pipelined loop (the CTAs of scan i+1 take over the SMs
# one by one as those of scan i exit): tests, then the loop at the 8-GPU shard size and at 10 M rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== tests"
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_scale.py -m gpu -q --tb=short -k "two_stream or graph or host or pipelined or scale" 2>&1 | tail -n 6
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k', 'host_enqueue_us_per_step')}, 'roofline', {k: d['roofline'][k] for k in ('frac', 'step_frac', 'kernel_ms')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
PY
}
for b in 1 32 64 128; do
  echo "== bench.py --rows 1250000 --batch $b (pipelined loop on one shard)"
  timeout 300 python bench.py --rows 1250000 --batch $b --steps 400 --warmup 10 --sweep 0 --check 0 --no-cpu > $O/r2_bench_shard_b$b.json 2> $O/r2_bench_shard.err; tail -c 300 $O/r2_bench_shard.err; show $O/r2_bench_shard_b$b.json
done
echo "== bench.py N = 1 (10 M rows)"
timeout 600 python bench.py --steps 100 --warmup 5 --sweep 0 --no-cpu > $O/r2_bench_n1_streams.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err; show $O/r2_bench_n1_streams.json
echo "== done"
