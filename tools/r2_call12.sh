#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== launch list of the timed steps (eager)"
  SCP_BENCH_CUDA_PROFILER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_timed.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-graph > gpurun_out/bench_under_ncu.log 2>&1
  echo "launch list rc=$?"; wc -l gpurun_out/launches_timed.csv; gzip -f gpurun_out/launches_timed.csv
  echo "== bench default (full line)"
  timeout 900 python bench.py > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; cut -c1-400 gpurun_out/r2_bench_line.json
  echo "== bench reference arm"
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; cut -c1-600 gpurun_out/r2_bench_reference_arm.json
  echo "== bench config1"
  timeout 600 python bench.py --workload config1 > gpurun_out/r2_bench_config1.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_config1.json
  echo "== bench bf16 fast mode (labelled)"
  timeout 600 python bench.py --vit-precision bf16 --no-cpu-baseline --no-kernel-breakdown > gpurun_out/r2_bench_bf16_fast_mode.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_bf16_fast_mode.json
} 2>&1 | tee gpurun_out/r2_call12.log
