#!/bin/bash
# Round-2 GPU call 22 (TWO B200s): the multi-GPU parity test (now with a hybrid search of limit 15 on the sharded
# index) and the host-buffer loop at the north-star shard size with 2, 3 and 4 batches in flight.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== multi-GPU parity test (world 2)"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short 2>&1 | tail -n 30 | tee $O/r2_pytest_gpu_multi_n2.log
echo "== ANN plugin with limit > 128"
timeout 300 python -m pytest tests/test_gpu_embeddings.py -m gpu -q --tb=short -k "ann_contract" 2>&1 | tail -n 8
for f in 2 3 4; do
  echo "== bench.py N = 2, 2.5 M rows (1.25 M per GPU), --inflight $f"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520 + f)) \
    bench.py --gpus 2 --rows 2500000 --steps 400 --warmup 10 --sweep 0 --check 0 --no-cpu --inflight $f > $O/r2_bench_n2_shard_f$f.json 2> $O/r2_bench_n2_shard.err; tail -c 300 $O/r2_bench_n2_shard.err | grep -v OMP | tail -n 3
  python - $O/r2_bench_n2_shard_f$f.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'host_enqueue_us_per_step')}, 'e2e', {k: d['e2e'][k] for k in ('value', 'ms_per_step', 'in_flight', 'host_enqueue_us_per_step', 'one_at_a_time_ms_per_step')})
PY
done
echo "== done"
