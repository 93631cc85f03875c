#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== full GPU suite"
  timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
  echo "== ncu evidence"
  bash tools/gpu_ncu.sh 2>&1 | tail -12
} 2>&1 | tee gpurun_out/r2_call10.log
