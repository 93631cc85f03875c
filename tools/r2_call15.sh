#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== ViT tests"
  timeout 900 python -m pytest tests/test_vit_gpu.py -m gpu -q 2>&1 | tail -3
  echo "== ViT timing"
  timeout 300 python tools/time_vit.py 64 2>&1 | tail -1
  echo "== corr / model tests"
  timeout 900 python -m pytest tests/test_corr_gpu.py tests/test_model_gpu.py tests/test_hotpath_gpu.py -m gpu -q 2>&1 | tail -3
  echo "== bench trainer"
  timeout 600 python bench.py --no-cpu-baseline 2>/dev/null > gpurun_out/r2_bench_line_b.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_line_b.json')); print({k: d.get(k) for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['hotpath']['ms_per_step'], d['config']['cuda_graph'])
for k in d['kernels']: print('%-70s %8.3f ms x%d  %8.1f %s frac %.4f' % (k['kernel'][:70], k['ms'], k['launches_per_step'], k['achieved'], k['unit'], k['frac']))"
} 2>&1 | tee gpurun_out/r2_call15.log
