#!/bin/bash
# Round-2 GPU call 1: R-GPU parity + legacy timing, Trainer.step timing, pose fit on GPU, prepared candidates A/B.
set -u
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
  echo "== R-GPU parity"
  timeout 600 python -m pytest tests/test_rgpu_softras_gpu.py -m gpu -q -s 2>&1 | grep -e PARITY -e passed -e failed -e Error | tail -40
  echo "== R-GPU timing (B=64, 256, uv1280)"
  timeout 600 python tools/time_rgpu.py 64 256 uv1280 2>&1 | tail -8
  echo "== Trainer.step timing"
  timeout 600 python tools/time_trainer.py 64 2>&1 | tail -3
  echo "== pose fit (first GPU run)"
  timeout 300 python -m pytest tests/test_posefit.py -m gpu -q -rxX 2>&1 | tail -4
  timeout 300 python tools/time_posefit.py 2>&1 | tail -5
  echo "== attention: early S issue"
  timeout 600 python tools/ab_variants.py --only early_qk --tests tests/test_vit_gpu.py --time "tools/time_vit.py 64"
  echo "== SoftRas candidates"
  timeout 1200 python tools/ab_variants.py --only fwd2px,facesmem,face16x2,softras_all --tests "tests/test_softras_gpu.py -k uv1280" \
      --time "tools/time_softras.py 64 256 uv1280"
} 2>&1 | tee gpurun_out/r2_call1.log
