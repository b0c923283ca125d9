#!/bin/bash
# round 2, first GPU call: correctness of the cluster kernel, then sweeps on the three big shapes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout 900 python -m pytest tests/test_gpu_gfpush.py -x -q -m gpu -k "tiers or cluster or surfaces or two_streams or tiny or golden" > gpurun_out/r02a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
export SWEEP_STEPS=4
timeout 600 python tools/sweep_gfpush.py reddit "push_cluster=0" "push_cluster=-1" "push_cluster=2" "push_cluster=4" "push_cluster=2,push_hub_deg=100000" "push_cluster=-1,push_cluster_probe=2" > gpurun_out/r02a_sweep_reddit.log 2>&1
tail -20 gpurun_out/r02a_sweep_reddit.log
timeout 900 python tools/sweep_gfpush.py mag "push_cluster=0" "push_cluster=2" "push_cluster=4" > gpurun_out/r02a_sweep_mag.log 2>&1
tail -12 gpurun_out/r02a_sweep_mag.log
SWEEP_SOURCES=4096 timeout 900 python tools/sweep_gfpush.py amazon2m "push_cluster=0" "push_cluster=16" "push_cluster=8" "push_cluster=16,push_hub_deg=256" > gpurun_out/r02a_sweep_amazon.log 2>&1
tail -14 gpurun_out/r02a_sweep_amazon.log
