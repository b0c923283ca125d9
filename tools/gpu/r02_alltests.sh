#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02t}
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_alltests.log 2>&1
echo "alltests rc=$?" >> gpurun_out/${TAG}_alltests.log
tail -6 gpurun_out/${TAG}_alltests.log
