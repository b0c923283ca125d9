#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export SWEEP_STEPS=4
for R in 1e-4 1e-3 3e-5; do
echo "rmax $R"
SWEEP_RMAX=$R timeout 300 python tools/sweep_gfpush.py reddit "push_cluster=0" "push_bucket=0" "push_bucket_block=256" "push_bucket_block=1024" 2>&1 | grep -A1 "rows/s" | cut -c1-170
done
