#!/bin/bash
# last call of the round: the rebuilt library at the final commit -- smoke + the tests of the last session's kernels
set -u
mkdir -p gpurun_out
{
  timeout 100 python -m pytest tests/test_corr_gpu.py tests/test_posefit.py -m gpu -q 2>&1 | tail -2
  timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_call42.log
