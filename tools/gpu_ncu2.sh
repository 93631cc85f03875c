#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_gpu.py tests/test_hotpath_gpu.py -x -q -s 2>&1 | grep -E "PARITY dino_argmatch|PARITY hotpath|passed|failed|Error|error|assert" | cut -c1-300 | head -20
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"forward_kernel|backward_kernel|pack_kernel|layernorm|depth_sums|project_faces|spmm3|im2col" -o /tmp/step2 python tools/ncu_targets.py > gpurun_out/ncu_step2.log 2>&1; echo "full rc=$?"
ncu -i /tmp/step2.ncu-rep --page raw --csv > gpurun_out/step_raw2.csv 2>/dev/null; ls -la gpurun_out/step_raw2.csv
ncu -i /tmp/step2.ncu-rep --page source --csv --kernel-name backward_kernel --launch-count 1 > gpurun_out/softras_bwd_source.csv 2>gpurun_out/src_err.log
gzip -f gpurun_out/softras_bwd_source.csv
ls -la gpurun_out/ | tail -6
