#!/bin/bash
# end-of-round validation on one GPU: every -m gpu test, smoke(), the default bench line, the reference arm, sanitizers
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${TAG:-r02f}
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_alltests.log 2>&1; echo "alltests rc=$?" >> gpurun_out/${TAG}_alltests.log
tail -3 gpurun_out/${TAG}_alltests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 400 gpurun_out/${TAG}_bench_ref.json; echo
if [ -z "$NO_SANITIZE" ]; then
  for TOOL in memcheck racecheck; do
    timeout 1500 compute-sanitizer --tool $TOOL python tools/sanitize_gfpush.py > gpurun_out/${TAG}_sanitize_${TOOL}.log 2>&1
    tail -4 gpurun_out/${TAG}_sanitize_${TOOL}.log
  done
fi
