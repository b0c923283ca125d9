#!/bin/bash
# every --workload of bench.py runs and prints one contract line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for WL in cora pubmed small amazon2m mag; do
  timeout 600 python bench.py --workload $WL --steps 3 --warmup 3 --no-side --no-cpu > gpurun_out/r02_wl_${WL}.json 2> gpurun_out/r02_wl_${WL}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_wl_${WL}.json"))
    r=d["roofline_gfpush"]
    print("${WL}", round(d["value"]), "rows/s  e2e", round(d["e2e"]["value"]), " kernel", r["kernel"], "frac", round(r["frac"],3), "bucket", r["bucket_count"], "slots", r["table_slots"], "agg frac", round(d["roofline_aggregate"]["frac"],2))
except Exception as e:
    print("${WL} FAILED", e); print(open("gpurun_out/r02_wl_${WL}.err").read()[-1500:])
PY
done
