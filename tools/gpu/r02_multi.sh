#!/bin/bash
# N-GPU run (gpurun --gpus N): the NCCL parity test of the sharded entry point, then the default bench line at N ranks
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${N:-2}
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multigpu.py -q -m gpu > gpurun_out/r02_multi${N}_tests.log 2>&1; tail -2 gpurun_out/r02_multi${N}_tests.log; fi
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -c 1500 gpurun_out/r02_bench_${N}gpu.json
