#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02ab}
timeout 900 python -m pytest tests/test_gpu_gfpush.py -q -m gpu -k "bucket" > gpurun_out/${TAG}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
export SWEEP_STEPS=4
B="push_bucket=2,push_bucket_block=512"
timeout 300 python tools/sweep_gfpush.py reddit "$B,push_bucket_fill=4" "$B,push_bucket_fill=5" "$B,push_bucket_fill=6" "$B,push_bucket_fill=7" "$B,push_bucket_fill=6,push_bucket_nb=8" > gpurun_out/${TAG}_sweep_reddit.log 2>&1
grep -A1 "rows/s" gpurun_out/${TAG}_sweep_reddit.log | cut -c1-200
timeout 400 python tools/sweep_gfpush.py mag "$B,push_bucket_fill=4" "$B,push_bucket_fill=5" "$B,push_bucket_fill=6" "$B,push_bucket_fill=7" > gpurun_out/${TAG}_sweep_mag.log 2>&1
grep -A1 "rows/s" gpurun_out/${TAG}_sweep_mag.log | cut -c1-200
SWEEP_SOURCES=4096 timeout 600 python tools/sweep_gfpush.py amazon2m "push_bucket_fill=4" "push_bucket_fill=5" "push_bucket_fill=6" > gpurun_out/${TAG}_sweep_amazon.log 2>&1
grep -A1 "rows/s" gpurun_out/${TAG}_sweep_amazon.log | cut -c1-200
