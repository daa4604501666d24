#!/bin/bash
# Round-2 GPU call 19 (TWO B200s): the multi-GPU parity test with k > 128, and the pipelined loops at the north-star
# SHARD size (2 x 1.25 M rows) with the host enqueue time per step: is the host-buffer loop host-bound at 0.29 ms?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== multi-GPU parity test (world 2)"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short 2>&1 | tail -n 30 | tee $O/r2_pytest_gpu_multi_n2.log
echo "== bench.py N = 2, 2.5 M rows (1.25 M per GPU)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --rows 2500000 --steps 300 --warmup 10 --sweep 0 --check 0 --no-cpu > $O/r2_bench_n2_shard.json 2> $O/r2_bench_n2_shard.err; tail -c 600 $O/r2_bench_n2_shard.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_bench_n2_shard.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms', 'host_enqueue_us_per_step')}, d['config']['exchange'])
print('e2e', d['e2e'])
PY
echo "== done"
