#!/bin/bash
# Round-2 GPU call 26 (EIGHT B200s), the record of the final build: the multi-GPU parity test on 8 ranks, the N = 8
# bench line (north-star operating point) and BASELINE configs[3] at scale (100 M x 1024 fp16, B = 64, top-100).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== multi-GPU parity test (world 8)"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -k "8" 2>&1 | tail -n 30 | tee $O/r2_pytest_gpu_multi_n8.log
run() {  # name, nproc, extra args...
  local name=$1 np=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + np)) \
    bench.py --gpus $np "$@" > $O/$name.json 2> $O/$name.err; tail -c 400 $O/$name.err | grep -v OMP_NUM_THREADS | tail -n 5
  python - "$O/$name.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
except Exception as e:
    print(sys.argv[1], 'unreadable', e); sys.exit(0)
print({k: d[k] for k in ('metric', 'n_gpus', 'value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k', 'host_enqueue_us_per_step')}, d['config']['exchange'])
print('roofline', {k: d['roofline'][k] for k in ('frac', 'step_frac', 'kernel_ms', 'kernel')}, 'e2e', {k: d['e2e'][k] for k in ('value', 'ms_per_step', 'in_flight', 'host_enqueue_us_per_step', 'one_at_a_time_ms_per_step')}, d['clocks'])
print('independent', d['independent_check']); print('sharded_equals_single', d['sharded_equals_single'])
for r in d['sweep'] or []: print(r['batch'], round(r['ms'], 4), round(r['scan_ms'], 4), round(r['hbm_frac'], 3), round(r['tensor_frac'], 3), r['family'][:34])
PY
}
echo "== bench.py N = 8 (10 M x 768 bf16 row-sharded, B = 32, top-10)"
run r2_bench_n8 8 --steps 200 --warmup 10
echo "== bench.py N = 8, --config D (100 M x 1024 fp16, B = 64, top-100)"
run r2_bench_cfgd_n8 8 --config D --steps 20 --warmup 3 --sweep 0
echo "== done"
