#!/bin/bash
# final validation of the round: full GPU suite, smoke, the bench lines, the launch list of one step and ncu --set full of the new
# correspondence kernels
set -u
mkdir -p gpurun_out
{
  echo "== full GPU suite"
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  echo "== bench default (full line)"
  timeout 900 python bench.py > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_line.json')); print({k: d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['hotpath']); print(d['roofline']); print(d['cpu_baseline'])
for k in d['kernels']: print('%-70s %8.3f ms x%d  %8.1f %s frac %.4f' % (k['kernel'][:70], k['ms'], k['launches_per_step'], k['achieved'], k['unit'], k['frac']))"
  echo "== bench reference arm"
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
  echo "== bench config1"
  timeout 300 python bench.py --workload config1 --no-cpu-baseline > gpurun_out/r2_bench_config1.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_config1.json
  echo "== launch list of one eager step"
  timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python tools/ncu_targets.py > gpurun_out/launches_step.log 2>&1
  echo "rc=$? lines=$(wc -l < gpurun_out/launches_step.csv)"; gzip -f gpurun_out/launches_step.csv
  echo "== ncu --set full of the correspondence GEMMs"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_rowstats_kernel" -c 2 -o /tmp/corr_full python tools/ncu_corr.py > gpurun_out/corr_full.log 2>&1
  echo "rc=$?"
  ncu -i /tmp/corr_full.ncu-rep --page raw --csv > gpurun_out/corr_full_raw.csv 2>/dev/null
  ncu -i /tmp/corr_full.ncu-rep --page source --csv --launch-count 1 > gpurun_out/corr_rows_source.csv 2>/dev/null; gzip -f gpurun_out/corr_rows_source.csv
  ls -la gpurun_out | tail -5
} 2>&1 | tee gpurun_out/r2_call39.log
