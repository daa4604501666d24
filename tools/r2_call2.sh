#!/bin/bash
# Round-2 GPU call 2 (one B200): full GPU suite on the flipped defaults, smoke, per-CTA timeline of the headline
# kernel (where the fixed cost goes), the new bench line (N = 1) and config D's shard, ncu traffic at 10 M rows.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== full GPU test suite"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 | tee $O/r2_pytest_gpu_call2.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3
echo "== timeline of mma_topk_kernel at the 8-GPU shard size"
ROWS=1250000 BATCHES=1,16,32 timeout 300 python tools/timeline_probe.py 2>&1 | tail -n 1 | tee $O/r2_timeline_shard.json
ROWS=1250000 BATCHES=32 KNOBS=mma_kps=1 timeout 300 python tools/timeline_probe.py 2>&1 | tail -n 1 | tee $O/r2_timeline_shard_kps1.json
ROWS=10000000 BATCHES=32 timeout 300 python tools/timeline_probe.py 2>&1 | tail -n 1 | tee $O/r2_timeline_10m.json
echo "== bench.py N = 1"
timeout 900 python bench.py --steps 50 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 600 $O/r2_bench_n1.err; cut -c1-3000 $O/r2_bench_n1.json
echo "== bench.py --config D (one rank's 12.5 M x 1024 fp16 shard)"
timeout 600 python bench.py --config D --steps 10 --warmup 3 --no-cpu > $O/r2_bench_cfgd_n1.json 2> $O/r2_bench_cfgd.err; tail -c 600 $O/r2_bench_cfgd.err; cut -c1-2500 $O/r2_bench_cfgd_n1.json
echo "== reference arm (CPU, full 10 M-row steps)"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_reference_arm.json 2>&1; cut -c1-1200 $O/r2_bench_reference_arm.json
echo "== ncu: dram traffic of the headline launch at 10 M rows"
ROWS=10000000 K=10 MODE=fast BATCHES=32 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:mma_topk" -s 3 -c 1 \
    -f -o $O/r2_mma_b32_10m python tools/tune_worker.py > $O/r2_mma_b32_10m.log 2>&1
python tools/ncu_summary.py $O/r2_mma_b32_10m.ncu-rep 2>&1 | head -n 8
echo "== done"
