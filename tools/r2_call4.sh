#!/bin/bash
# Round-2 GPU call 4: fused jitter / symmetry parity, full GPU suite, Trainer.step timing + profile.
set -u
mkdir -p gpurun_out
{
  echo "== new-kernel tests"
  timeout 600 python -m pytest tests/test_jitter_gpu.py tests/test_symmetry_gpu.py -m gpu -q -s 2>&1 | grep -e PARITY -e passed -e failed -e Error -e error -e assert | tail -40
  echo "== full GPU suite"
  timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
  echo "== Trainer.step timing"
  timeout 600 python tools/time_trainer.py 64 2>&1 | tail -2
  echo "== Trainer.step profile"
  timeout 600 python tools/profile_trainer.py 64 2>&1 | grep -v Warn | head -120
} 2>&1 | tee gpurun_out/r2_call4.log
