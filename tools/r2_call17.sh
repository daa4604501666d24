#!/bin/bash
# Round-2 GPU call 17 (one B200): K1 bulk-copy kernel under ncu (kernel duration, DRAM bytes) + host launch cost.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== ncu"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,launch__cluster_size --clock-control none -k regex:pool --csv --log-file gpurun_out/r2_k1_bulk_launches.csv python tools/k1_probe.py > gpurun_out/r2_k1_ncu.err 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2_k1_bulk_launches.csv")) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = {}
for r in rows[1:]:
    key = (r[ix["ID"]], r[ix["Kernel Name"]][:60], r[ix["Grid Size"]] if "Grid Size" in ix else "")
    agg.setdefault(key, {})[r[ix["Metric Name"]]] = r[ix["Metric Value"]] + " " + r[ix["Metric Unit"]]
for k, v in list(agg.items()):
    print(k, v)
PY
echo "== host launch cost"
timeout 300 python tools/k1_probe.py host
echo "== done"
