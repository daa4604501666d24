#!/bin/bash
# Round-2 GPU call 38 (one B200): the GPU suite minus tests/test_gpu_experimental.py (run in full in call 33, its
# screen-mode additions in call 37) on the final tree, and smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short --ignore=tests/test_gpu_experimental.py 2>&1 | tail -n 8
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -n 2
echo "== done"
