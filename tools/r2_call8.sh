#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== graphed step test + model tests"
  timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_jitter_gpu.py tests/test_symmetry_gpu.py -m gpu -q -s -x 2>&1 | grep -e PARITY -e passed -e failed -e Error -e "^E " | tail -30 | cut -c1-1500
  echo "== bench (trainer, graph)"
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_trainer_graph.json 2> gpurun_out/r2_bench_trainer_graph.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_trainer_graph.json'))
    print({k: d[k] for k in ('value','ms_per_step','e2e','hotpath','loss')}, d['config']['cuda_graph'])
except Exception as e:
    print('ERR', e)
PY
  grep -v "Warn\|warn" gpurun_out/r2_bench_trainer_graph.err | tail -12
} 2>&1 | tee gpurun_out/r2_call8.log
