#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== pose fit tests"
  timeout 300 python -m pytest tests/test_posefit.py -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|Error|assert" | tail -8
  echo "== pose fit timing"
  timeout 200 python tools/time_posefit.py 2>&1 | tail -4
} 2>&1 | tee gpurun_out/r2_call37.log
