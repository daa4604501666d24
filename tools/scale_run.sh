#!/bin/bash
# 1 -> 8 GPU scaling run of bench.py (same launch line the driver uses)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps ${STEPS:-200} --warmup 10 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps ${STEPS:-200} --warmup 10 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","recall_at_10")}, "roof", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), d["clocks"])
    print([(s["batch"], round(s.get("ms",0),3), round(s.get("hbm_frac",0),3)) for s in d["sweep"]])
except Exception as e:
    print("parse fail", e); print(open("gpurun_out/scale_n$N.err").read()[-1500:])
PY
done
