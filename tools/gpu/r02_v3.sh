#!/bin/bash
# bucket kernel at 512 threads: parity of every bucket tier, then a per-line ncu profile on the Reddit-shape graph
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out


TAG=r02x_b512 WL=reddit CFG="push_bucket=2,push_bucket_block=512" SRC=8192 KREGEX=gfpush_bucket_kernel bash tools/gpu/r02_prof.sh
python tools/ncu_lines.py gpurun_out/prof_r02x_b512.ncu-rep gfpush_bucket_kernelILi512E 400 > gpurun_out/r02x_b512_lines.txt 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02x_b512.ncu-rep "r02 bucket kernel, 512 threads x 2 CTAs per SM, Reddit-shape, 8192 sources" > gpurun_out/r02x_b512_summary.md 2>/dev/null
rm -f gpurun_out/prof_r02x_b512.ncu-rep
