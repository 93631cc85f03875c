#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== full GPU suite"
  timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
  echo "== smoke"
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  echo "== bench default (full line)"
  timeout 900 python bench.py > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_line.json')); print({k: d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['hotpath']); print(d['roofline']); print(d['cpu_baseline']); print(d.get('data_path'))
for k in d['kernels']: print('%-70s %8.3f ms x%d  %8.1f %s frac %.4f' % (k['kernel'][:70], k['ms'], k['launches_per_step'], k['achieved'], k['unit'], k['frac']))"
  echo "== bench reference arm"
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
  echo "== bench config1"
  timeout 600 python bench.py --workload config1 --no-cpu-baseline > gpurun_out/r2_bench_config1.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_config1.json
  echo "== bench bf16 fast mode"
  timeout 600 python bench.py --vit-precision bf16 --no-cpu-baseline --no-kernel-breakdown > gpurun_out/r2_bench_bf16.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_bf16.json
  echo "== ncu"
  bash tools/gpu_ncu.sh 2>&1 | tail -12
} 2>&1 | tee gpurun_out/r2_call29.log
