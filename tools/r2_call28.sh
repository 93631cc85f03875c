#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== nhwc glue tests"
  timeout 600 python -m pytest tests/test_nhwc_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -12
  echo "== bench"
  timeout 900 python bench.py --no-cpu-baseline --no-kernel-breakdown > gpurun_out/r2_bench_nhwc.json 2> gpurun_out/r2_bench_nhwc.err; tail -3 gpurun_out/r2_bench_nhwc.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_nhwc.json')); print({k: d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})"
} 2>&1 | tee gpurun_out/r2_call28.log
