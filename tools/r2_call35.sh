#!/bin/bash
# Round-2 GPU call 35 (one B200): sustained B = 32 at 10 M rows (power-capped) with the screen-mode knob: half the
# tensor work in the scan, exact re-scoring in the reduce (ss_screen), with and without shared memory left for the
# re-scoring reduce to co-reside (smem_reserve_kb).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k')}, 'roofline', {k: d['roofline'][k] for k in ('frac', 'step_frac', 'kernel_ms')}, 'e2e', d['e2e']['value'], d['clocks'])
PY
}
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 200 --warmup 10 --sweep 0 --check 0 --no-cpu > $O/r2_ss.json 2> $O/r2_ss.err; tail -c 200 $O/r2_ss.err; show $O/r2_ss.json; }
run VQA_SS_SCREEN=0
run VQA_SS_SCREEN=1
run VQA_SS_SCREEN=1 VQA_SMEM_RESERVE_KB=24
run VQA_SS_SCREEN=0
echo "== done"
