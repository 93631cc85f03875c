#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== parity tests"
  timeout 1500 python -m pytest tests/test_hotpath_gpu.py tests/test_model_gpu.py tests/test_posefit.py -m gpu -q -s 2>&1 | grep -e PARITY -e passed -e failed -e Error -e "^E " | tail -40
  echo "== smoke"
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_call6.log
