#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== data path tests"
  timeout 600 python -m pytest tests/test_data_gpu.py -m gpu -q -x -s 2>&1 | grep -E "PARITY|passed|failed|Error|error|assert" | head -30
} 2>&1 | tee gpurun_out/r2_call26.log
