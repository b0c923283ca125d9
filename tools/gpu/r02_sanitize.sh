#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  SANITIZE_ONLY=${ONLY:-bucket} timeout 1500 compute-sanitizer --tool $TOOL python tools/sanitize_gfpush.py > gpurun_out/r02b_sanitize_${TOOL}.log 2>&1
  tail -6 gpurun_out/r02b_sanitize_${TOOL}.log
done
