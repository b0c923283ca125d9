#!/bin/bash
# bucket kernel iteration: parity of every bucket tier + sweeps on the three shapes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02y}
timeout 900 python -m pytest tests/test_gpu_gfpush.py -q -m gpu -k "bucket" > gpurun_out/${TAG}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
export SWEEP_STEPS=4
timeout 300 python tools/sweep_gfpush.py reddit "push_cluster=0" "push_bucket=2" "push_bucket=2,push_bucket_block=512" "push_bucket=2,push_bucket_block=256" > gpurun_out/${TAG}_sweep_reddit.log 2>&1
grep -A1 "rows/s" gpurun_out/${TAG}_sweep_reddit.log | cut -c1-200
timeout 400 python tools/sweep_gfpush.py mag "push_cluster=0" "push_bucket=2,push_bucket_block=512" "push_bucket=2,push_bucket_block=256" > gpurun_out/${TAG}_sweep_mag.log 2>&1
grep -A1 "rows/s" gpurun_out/${TAG}_sweep_mag.log | cut -c1-200
SWEEP_SOURCES=4096 timeout 600 python tools/sweep_gfpush.py amazon2m "push_bucket=1" "push_bucket_block=512" > gpurun_out/${TAG}_sweep_amazon.log 2>&1
grep -A1 "rows/s" gpurun_out/${TAG}_sweep_amazon.log | cut -c1-200
