#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== model tests (tail graph)"
  timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -15
  echo "== bench default"
  timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_tail.json 2> gpurun_out/r2_bench_tail.err; tail -3 gpurun_out/r2_bench_tail.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_tail.json')); print({k: d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})"
} 2>&1 | tee gpurun_out/r2_call22.log
