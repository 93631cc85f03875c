#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== corr tcgen05 forward: parity"
  timeout 240 python -m pytest tests/test_corr_gpu.py -m gpu -q -s -x -k "tc_forward or rotation or golden" 2>&1 | grep -v "^$" | tail -25
  echo "== corr forward timing"
  timeout 150 python tools/time_corr.py 2>&1 | tail -8
} 2>&1 | tee gpurun_out/r2_call32.log
