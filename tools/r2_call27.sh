#!/bin/bash
# Round-2 GPU calls 27 / 28 (FOUR, then TWO B200s): the N = 4 and N = 2 bench lines on the final build.
# usage: bash tools/r2_call27.sh <n_gpus>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
np=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + np)) \
  bench.py --gpus $np --steps 200 --warmup 10 --sweep 0 > $O/r2_bench_n$np.json 2> $O/r2_bench_n$np.err; tail -c 400 $O/r2_bench_n$np.err | grep -v OMP_NUM_THREADS | tail -n 5
python - "$O/r2_bench_n$np.json" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: d[k] for k in ('metric', 'n_gpus', 'value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k', 'host_enqueue_us_per_step')}, d['config']['exchange'])
print('roofline', {k: d['roofline'][k] for k in ('frac', 'step_frac', 'kernel_ms', 'kernel')}, 'e2e', {k: d['e2e'][k] for k in ('value', 'ms_per_step', 'in_flight', 'one_at_a_time_ms_per_step')}, d['clocks'])
print('independent', d['independent_check']['ids_equal_independent'], d['independent_check']['fast_vs_independent']); print('sharded_equals_single', d['sharded_equals_single'])
PY
echo "== done"
