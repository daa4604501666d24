#!/bin/bash
# Round-2 GPU call 14 (one B200): does leaving shared memory for the re-scoring reduce (VQA_SMEM_RESERVE_KB) shorten the
# pipelined loop for 33..256 queries? Shard size (1.25 M rows) and 10 M rows, reserve 0 (default) against 24 KB.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k')}, 'roofline', {k: d['roofline'][k] for k in ('frac', 'step_frac', 'kernel_ms')}, 'kernel', d['roofline'].get('kernel'))
PY
}
for rows in 1250000 10000000; do
  steps=200; [ $rows = 10000000 ] && steps=40
  for b in 64 128 256; do
    for r in 0 24; do
      echo "== rows $rows batch $b VQA_SMEM_RESERVE_KB=$r"
      VQA_SMEM_RESERVE_KB=$r timeout 300 python bench.py --rows $rows --batch $b --steps $steps --warmup 5 --sweep 0 --check 0 --no-cpu > $O/r2_rsv_${rows}_b${b}_r$r.json 2> $O/r2_rsv.err; tail -c 300 $O/r2_rsv.err; show $O/r2_rsv_${rows}_b${b}_r$r.json
    done
  done
done
echo "== done"
