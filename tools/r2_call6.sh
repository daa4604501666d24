#!/bin/bash
# Round-2 GPU call 6 (TWO B200s): row-sharded search -- the multi-GPU parity test (nccl and p2p exchange, pipelined
# and host-buffer forms, Embeddings(shards=True)) and the N = 2 bench line with both exchanges.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== multi-GPU parity test (world 2)"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short 2>&1 | tail -n 30 | tee $O/r2_pytest_gpu_multi_n2.log
echo "== bench.py N = 2 (auto exchange = peer-memory push when symmetric memory is available)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 50 --warmup 5 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err; tail -c 600 $O/r2_bench_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n2.json',):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print({k:d[k] for k in ('value','ms_per_step','one_step_at_a_time_ms','recall_at_k')}, d['config']['exchange'], d['roofline']['frac'], d['roofline']['step_frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['one_at_a_time_ms_per_step'])
    print(d['independent_check']); print(d['sharded_equals_single'])
    for r in d['sweep']: print(r['batch'], round(r['ms'],4), round(r['scan_ms'],4), round(r['hbm_frac'],3), r['family'][:30])
PY
echo "== bench.py N = 2, NCCL exchange"
VQA_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 50 --warmup 5 --sweep 0 --check 0 > $O/r2_bench_n2_nccl.json 2> $O/r2_bench_n2_nccl.err; tail -c 300 $O/r2_bench_n2_nccl.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_n2_nccl.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','one_step_at_a_time_ms')}, d['config']['exchange'], d['e2e']['value'])
except Exception as e:
    print('unreadable', e)
PY
echo "== done"
