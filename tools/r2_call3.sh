#!/bin/bash
# Round-2 GPU call 3: first run of the fp32-class (x3) ViT path: unit tests, timing; Trainer.step profile by operator.
set -u
mkdir -p gpurun_out
{
  echo "== ViT tests (x3 + bf16)"
  timeout 900 python -m pytest tests/test_vit_gpu.py tests/test_symmetry_gpu.py -m gpu -q -s 2>&1 | grep -e PARITY -e passed -e failed -e Error -e error -e assert | tail -80
  echo "== ViT timing"
  timeout 300 python tools/time_vit.py 64 2>&1 | tail -3
  echo "== Trainer.step profile"
  timeout 600 python tools/profile_trainer.py 64 2>&1 | tail -150
} 2>&1 | tee gpurun_out/r2_call3.log
