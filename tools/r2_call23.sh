#!/bin/bash
# Round-2 GPU call 23 (one B200): the full synccheck report (the filtered one of calls 9 / 21 kept only stack frames).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/r2_synccheck_full.txt 2>&1
grep -v "Host Frame" gpurun_out/r2_synccheck_full.txt | head -n 120
echo "== summary"; grep -c "Barrier error\|Error:" gpurun_out/r2_synccheck_full.txt; grep "ERROR SUMMARY" gpurun_out/r2_synccheck_full.txt
