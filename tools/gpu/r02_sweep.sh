#!/bin/bash
# sweeps only: reddit / mag / amazon2m with the default cluster sizes next to the per-CTA kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02s}
export SWEEP_STEPS=${SWEEP_STEPS:-8}
timeout 600 python tools/sweep_gfpush.py reddit ${REDDIT_CFGS:-"push_cluster=0" "push_cluster=2"} > gpurun_out/${TAG}_sweep_reddit.log 2>&1
grep -v wide_ gpurun_out/${TAG}_sweep_reddit.log | tail -8
timeout 900 python tools/sweep_gfpush.py mag ${MAG_CFGS:-"push_cluster=0" "push_cluster=2"} > gpurun_out/${TAG}_sweep_mag.log 2>&1
grep -v wide_ gpurun_out/${TAG}_sweep_mag.log | tail -6
SWEEP_STEPS=3 SWEEP_SOURCES=4096 timeout 900 python tools/sweep_gfpush.py amazon2m ${AMAZON_CFGS:-"push_cluster=0" "push_cluster=16"} > gpurun_out/${TAG}_sweep_amazon.log 2>&1
grep -v wide_ gpurun_out/${TAG}_sweep_amazon.log | tail -6
