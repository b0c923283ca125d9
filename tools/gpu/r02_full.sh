#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02e}
timeout 900 python -m pytest tests/test_gpu_gfpush.py -q -m gpu -k "tiers or cluster" > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "tiers_agree" > gpurun_out/${TAG}_full.log 2>&1
echo "full rc=$?" >> gpurun_out/${TAG}_full.log
tail -30 gpurun_out/${TAG}_full.log
SWEEP_STEPS=2 timeout 600 python tools/sweep_gfpush.py reddit "push_cluster=0" "push_cluster=2" > gpurun_out/${TAG}_sweep_reddit.log 2>&1
tail -6 gpurun_out/${TAG}_sweep_reddit.log
