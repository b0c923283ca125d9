#!/bin/bash
# validation of the rebuilt tree + block-size / forced-bucket sweeps on the table-sized shapes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NO_SANITIZE=1 TAG=r02v bash tools/gpu/r02_final.sh
export SWEEP_STEPS=4
timeout 300 python tools/sweep_gfpush.py reddit "push_cluster=0" "push_bucket=2" "push_bucket=2,push_bucket_nb=4" > gpurun_out/r02v_sweep_reddit.log 2>&1
SWEEP_BLOCK=512 timeout 300 python tools/sweep_gfpush.py reddit "push_smem_probe=2" "push_smem_probe=4" > gpurun_out/r02v_sweep_reddit_b512.log 2>&1
tail -8 gpurun_out/r02v_sweep_reddit.log gpurun_out/r02v_sweep_reddit_b512.log
timeout 400 python tools/sweep_gfpush.py mag "push_cluster=0" "push_bucket=2" > gpurun_out/r02v_sweep_mag.log 2>&1
tail -6 gpurun_out/r02v_sweep_mag.log
