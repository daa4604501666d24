#!/bin/bash
# Round-2 GPU call 29 (one B200): the full racecheck report over every kernel family, aggregated by location.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --print-limit 100000 python tools/sanitize.py > gpurun_out/r2_racecheck_full.txt 2>&1
echo "== racecheck hazards by kernel and source line"
grep "Race reported\|and Write access\|and Read access" gpurun_out/r2_racecheck_full.txt | sed 's/+0x[0-9a-f]*//g; s/\[[0-9]* hazards\]//' | sed 's/.*access at //' | cut -c1-160 | sort | uniq -c | sort -rn | head -n 30
grep "RACECHECK SUMMARY\|ERROR SUMMARY\|sanitize run complete" gpurun_out/r2_racecheck_full.txt
