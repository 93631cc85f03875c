#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
gzip -f gpurun_out/launches.csv
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"backward_face_kernel|forward_kernel|corr_bwd_rows_kernel|corr_fwd_kernel|gemm_bf16_tn_kernel|fa2_fwd_kernel" -o /tmp/step3 python tools/ncu_targets.py > gpurun_out/ncu_step3.log 2>&1; echo "full rc=$?"
ncu -i /tmp/step3.ncu-rep --page raw --csv > gpurun_out/step_raw3.csv 2>/dev/null; ls -la gpurun_out/step_raw3.csv
ncu -i /tmp/step3.ncu-rep --page source --csv --kernel-name backward_face_kernel --launch-count 1 > gpurun_out/softras_bwd_face_source.csv 2>/dev/null
gzip -f gpurun_out/softras_bwd_face_source.csv
