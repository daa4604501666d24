#!/bin/bash
# Round-2 GPU call 33 (one B200): the whole GPU suite and smoke on the final tree (after the mbarrier hand-over in the
# TS kernel, k / 64 segments for k > 128 and the host-call guard of the ANN plugin).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full GPU test suite"
timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -n 12 | tee gpurun_out/r2_pytest_gpu_final.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3
echo "== done"
