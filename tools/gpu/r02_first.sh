#!/bin/bash
# round 2: correctness of the cluster kernel, then sweeps on the three big shapes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02b}
timeout 900 python -m pytest tests/test_gpu_gfpush.py -x -q -m gpu -k "tiers or cluster or surfaces or two_streams or tiny or golden" > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
export SWEEP_STEPS=4
timeout 600 python tools/sweep_gfpush.py reddit "push_cluster=0" "push_cluster=-1" "push_cluster=2" "push_cluster=4" > gpurun_out/${TAG}_sweep_reddit.log 2>&1
tail -20 gpurun_out/${TAG}_sweep_reddit.log
timeout 900 python tools/sweep_gfpush.py mag "push_cluster=0" "push_cluster=2" "push_cluster=4" > gpurun_out/${TAG}_sweep_mag.log 2>&1
tail -12 gpurun_out/${TAG}_sweep_mag.log
SWEEP_SOURCES=4096 timeout 900 python tools/sweep_gfpush.py amazon2m "push_cluster=0" "push_cluster=16" "push_cluster=16,push_hub_deg=256" > gpurun_out/${TAG}_sweep_amazon.log 2>&1
tail -14 gpurun_out/${TAG}_sweep_amazon.log
