#!/bin/bash
# ncu evidence of one round: launch list of the bench command + --set full of one real step (raw CSV only: the
# .ncu-rep stays on the box, gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
gzip -f gpurun_out/launches.csv
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"softras|corr|gemm|vit|fa2|loss|geom|cycle" -o /tmp/step python tools/ncu_targets.py > gpurun_out/ncu_step.log 2>&1; echo "full rc=$?"
ncu -i /tmp/step.ncu-rep --page raw --csv > gpurun_out/step_raw.csv 2>/dev/null; ls -la gpurun_out/step_raw.csv
ncu -i /tmp/step.ncu-rep --page source --csv --kernel-name regex:backward_kernel --launch-count 1 > gpurun_out/softras_bwd_source.csv 2>/dev/null
ncu -i /tmp/step.ncu-rep --page source --csv --kernel-name regex:fa2_fwd_kernel --launch-count 1 > gpurun_out/fa2_source.csv 2>/dev/null
gzip -f gpurun_out/softras_bwd_source.csv gpurun_out/fa2_source.csv
ls -la gpurun_out/ | tail -8
