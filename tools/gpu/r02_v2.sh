#!/bin/bash
# bucket kernel with two / three sources per SM: parity of the new tiers, then the sweeps
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gfpush.py -x -q -m gpu -k "bucket" > gpurun_out/r02w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_tests.log
tail -5 gpurun_out/r02w_tests.log
export SWEEP_STEPS=4
SWEEP_SOURCES=4096 timeout 600 python tools/sweep_gfpush.py amazon2m "push_bucket=1" "push_bucket_block=512" "push_bucket_block=256" "push_bucket_block=512,push_bucket_nb=64" > gpurun_out/r02w_sweep_amazon.log 2>&1
tail -12 gpurun_out/r02w_sweep_amazon.log
timeout 300 python tools/sweep_gfpush.py reddit "push_cluster=0" "push_bucket=2" "push_bucket=2,push_bucket_block=512" "push_bucket=2,push_bucket_block=256" "push_bucket=2,push_bucket_block=512,push_bucket_nb=4" "push_bucket=2,push_bucket_block=256,push_bucket_nb=8" > gpurun_out/r02w_sweep_reddit.log 2>&1
tail -14 gpurun_out/r02w_sweep_reddit.log
timeout 400 python tools/sweep_gfpush.py mag "push_cluster=0" "push_bucket=2,push_bucket_block=512" "push_bucket=2,push_bucket_block=256" > gpurun_out/r02w_sweep_mag.log 2>&1
tail -8 gpurun_out/r02w_sweep_mag.log
