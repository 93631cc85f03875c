#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== fa3 early-QK A/B"
  timeout 900 python tools/ab_variants.py --only fa3_late --tests "tests/test_vit_gpu.py -k x3" --time "tools/time_vit.py 64"
  echo "== rotation cycle test"
  timeout 300 python -m pytest tests/test_corr_gpu.py -m gpu -q 2>&1 | tail -2
  echo "== launch list of the timed steps (graph replays)"
  SCP_BENCH_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_timed.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown > gpurun_out/bench_under_ncu.log 2>&1
  echo "launch list rc=$?"; wc -l gpurun_out/launches_timed.csv; gzip -f gpurun_out/launches_timed.csv
  echo "== configs[4] sweep"
  timeout 600 python tools/sweep_kernels.py gpurun_out/r2_sweep.md 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_call11.log
