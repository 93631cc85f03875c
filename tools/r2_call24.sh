#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== ViT tests"
  timeout 900 python -m pytest tests/test_vit_gpu.py -m gpu -q -x 2>&1 | tail -5
  echo "== ViT timing B=64"
  timeout 600 python tools/time_vit.py 64 2>&1 | tail -12
} 2>&1 | tee gpurun_out/r2_call24.log
