#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== SoftRas face-centric backward: row-major bbox enumeration"
  timeout 900 python tools/ab_variants.py --only facelinear --tests tests/test_softras_gpu.py --time "tools/time_softras.py 64"
} 2>&1 | tee gpurun_out/r2_call23.log
