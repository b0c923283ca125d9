#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export SWEEP_STEPS=4
for WL in pubmed cora; do
timeout 300 python tools/sweep_gfpush.py $WL "push_cluster=0" > gpurun_out/r02ad_sweep_${WL}.log 2>&1
SWEEP_SCRATCH=2 timeout 300 python tools/sweep_gfpush.py $WL "push_bucket=2,push_bucket_block=1024" "push_bucket=2,push_bucket_block=512" "push_bucket=2,push_bucket_block=256" "push_bucket=0,push_smem_hash=2" >> gpurun_out/r02ad_sweep_${WL}.log 2>&1
grep -A1 "rows/s" gpurun_out/r02ad_sweep_${WL}.log | cut -c1-200
done
