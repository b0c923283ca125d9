#!/bin/bash
# ncu evidence for the bench command: full-set capture of the timed GFPush + aggregation kernels and the launch list.
# The reports are summarised ON THE BOX (gpurun copies back at most 64 MiB); only the Reddit-shape report is kept.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for WL in ${WLS:-reddit amazon2m mag}; do
  REP=gpurun_out/prof_r02_bench_${WL}
  # matched launches per step: gfpush_kernel + aggregate_fwd_kernel; where the bucket kernel runs first (amazon2m) also the slab
  # pass that takes its hand-overs, and one more bucket launch (the pilot) in the first warm-up step
  SKIP=6; COUNT=2; KSUB=gfpush_kernelILi1024E
  if [ "$WL" = "amazon2m" ]; then SKIP=10; COUNT=3; KSUB=gfpush_bucket_kernel; fi
  timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"gfpush_kernel|gfpush_bucket_kernel|aggregate_fwd_kernel" -s $SKIP -c $COUNT -f \
      -o ${REP} python bench.py --workload ${WL} --steps 2 --warmup 3 --no-side --no-cpu > gpurun_out/r02_ncu_${WL}.log 2>&1
  tail -1 gpurun_out/r02_ncu_${WL}.log
  ncu -i ${REP}.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${WL}_raw.csv 2>/dev/null
  python tools/ncu_lines.py ${REP}.ncu-rep $KSUB 40 > gpurun_out/r02_ncu_${WL}_lines.txt 2>&1
  python tools/ncu_summary.py ${REP}.ncu-rep "r02 bench ${WL}: GFPush kernel as launched by bench.py" > gpurun_out/r02_ncu_${WL}_gfpush.md 2>/dev/null
  if [ "$WL" != "reddit" ]; then rm -f ${REP}.ncu-rep; fi
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_${WL}.csv \
      python bench.py --workload ${WL} --steps 2 --warmup 3 --no-side --no-cpu > /dev/null 2>&1
  wc -l gpurun_out/r02_launches_${WL}.csv
done
du -sh gpurun_out
