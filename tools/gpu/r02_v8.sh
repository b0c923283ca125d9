#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02af}
timeout 900 python -m pytest tests/test_gpu_gfpush.py -q -m gpu -x -k "bucket or tiers or golden" > gpurun_out/${TAG}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
export SWEEP_STEPS=4
timeout 300 python tools/sweep_gfpush.py reddit "push_cluster=0" "push_bucket=0" > gpurun_out/${TAG}_sweep.log 2>&1
timeout 400 python tools/sweep_gfpush.py mag "push_cluster=0" >> gpurun_out/${TAG}_sweep.log 2>&1
SWEEP_SOURCES=4096 timeout 600 python tools/sweep_gfpush.py amazon2m "push_bucket=1" >> gpurun_out/${TAG}_sweep.log 2>&1
SWEEP_SCRATCH=2 timeout 300 python tools/sweep_gfpush.py pubmed "push_bucket=2,push_bucket_block=512" >> gpurun_out/${TAG}_sweep.log 2>&1
grep -A1 "rows/s" gpurun_out/${TAG}_sweep.log | cut -c1-200
