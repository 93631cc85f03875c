#!/bin/bash
set -u
mkdir -p gpurun_out
{
  for v in "" slim epi16 epi16slim; do
    echo "== variant '$v'"
    SCP_LIB_VARIANT=$v timeout 150 python tools/time_corr.py 2>&1 | tail -3
  done
} 2>&1 | tee gpurun_out/r2_call38.log
