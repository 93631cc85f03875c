#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== side-stream tests"
  timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "side_streams or graphed" 2>&1 | grep -e PARITY -e passed -e failed -e "^E " | cut -c1-400 | tail -8
  for ov in 1 0; do
    echo "== bench trainer, overlap_rotation=$ov"
    SCP_OVERLAP_ROTATION=$ov timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['config']['cuda_graph'])"
  done
  echo "== bench trainer, no ViT overlap"
  timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown --no-overlap 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['config']['cuda_graph'])"
} 2>&1 | tee gpurun_out/r2_call14.log
