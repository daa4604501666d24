#!/bin/bash
# Round-2 GPU call 24 (one B200): after removing the divergent loop exit in front of the TS kernel's block-wide barrier
# (M = 64 variant): synccheck over every kernel family (aggregated), the M = 64 / TS tests, config D on one shard.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 compute-sanitizer --tool synccheck --print-limit 100000 python tools/sanitize.py > $O/r2_synccheck_full.txt 2>&1
echo "== synccheck findings by kind and location"
grep "Barrier error\|Device Frame\| at vqa" $O/r2_synccheck_full.txt | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -n 20
grep "ERROR SUMMARY\|sanitize run complete" $O/r2_synccheck_full.txt
echo "== TS / M = 64 tests"
timeout 900 python -m pytest tests/test_gpu_experimental.py tests/test_gpu_search.py -m gpu -q --tb=short -k "m64 or ts or wide or fast_modes" 2>&1 | tail -n 6
echo "== bench.py --config D on one shard"
timeout 900 python bench.py --config D --steps 20 --warmup 3 --sweep 0 > $O/r2_bench_cfgd_n1.json 2> $O/r2_bench_cfgd_n1.err; tail -c 300 $O/r2_bench_cfgd_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_bench_cfgd_n1.json') if l.startswith('{')][-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'one_step_at_a_time_ms', 'recall_at_k')}, {k: d['roofline'].get(k) for k in ('frac', 'step_frac', 'kernel_ms')}, d['e2e']['value'], d['independent_check']['ids_equal_independent'])
PY
echo "== done"
