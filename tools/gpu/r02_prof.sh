#!/bin/bash
# ncu capture of one GFPush launch (full set, source view)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02c}; WL=${WL:-reddit}; CFG=${CFG:-push_cluster=2}; SRC=${SRC:-8192}; KREGEX=${KREGEX:-gfpush_cluster_kernel}
SWEEP_STEPS=1 SWEEP_SOURCES=$SRC timeout 900 ncu --set full --import-source on --clock-control none -k regex:$KREGEX -s 2 -c 1 -f -o gpurun_out/prof_${TAG} python tools/sweep_gfpush.py $WL "$CFG" > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
