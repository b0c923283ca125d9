#!/bin/bash
# ncu evidence for the bench command after the hash-bucket kernel became the default on every shape beyond the dense mode:
# full-set capture of the first timed step's launches (bucket kernel, slab pass for hand-overs, aggregation) and the launch list.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for WL in ${WLS:-reddit amazon2m mag}; do
  REP=gpurun_out/prof_r02b_bench_${WL}
  KSUB=gfpush_bucket_kernelILi512E; NLAUNCH=400; if [ "$WL" = "amazon2m" ]; then KSUB=gfpush_bucket_kernelILi1024E; fi; if [ "$WL" = "mag" ]; then NLAUNCH=1500; fi
  # matched launches: warm-up step 0 = pilot + bucket + slab + aggregate, later steps = bucket + slab + aggregate
  timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"gfpush_kernel|gfpush_bucket_kernel|aggregate_fwd_kernel" -s 10 -c 3 -f \
      -o ${REP} python bench.py --workload ${WL} --steps 2 --warmup 3 --no-side --no-cpu > gpurun_out/r02b_ncu_${WL}.log 2>&1
  tail -1 gpurun_out/r02b_ncu_${WL}.log
  ncu -i ${REP}.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_${WL}_raw.csv 2>/dev/null
  python tools/ncu_lines.py ${REP}.ncu-rep $KSUB 40 > gpurun_out/r02b_ncu_bench_${WL}_lines.txt 2>&1
  python tools/ncu_summary.py ${REP}.ncu-rep "r02b bench ${WL}: GFPush kernel as launched by bench.py (hash-bucket kernel, the default)" > gpurun_out/r02b_ncu_bench_${WL}_gfpush.md 2>/dev/null
  rm -f ${REP}.ncu-rep
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-400} --csv --log-file gpurun_out/r02b_launches_${WL}.csv \
      python bench.py --workload ${WL} --steps 2 --warmup 3 --no-side --no-cpu > /dev/null 2>&1
  wc -l gpurun_out/r02b_launches_${WL}.csv
done
du -sh gpurun_out
