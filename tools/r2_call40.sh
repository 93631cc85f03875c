#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== 2-GPU bench (final kernels)"
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-kernel-breakdown > gpurun_out/r2_bench_line_2gpu_final.json 2> gpurun_out/r2_bench_2gpu.err
  echo "rc=$?"; cut -c1-700 gpurun_out/r2_bench_line_2gpu_final.json; tail -3 gpurun_out/r2_bench_2gpu.err | cut -c1-300
} 2>&1 | tee gpurun_out/r2_call40.log
