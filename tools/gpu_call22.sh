#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vit_gpu.py tests/test_corr_gpu.py tests/test_losses_gpu.py tests/test_hotpath_gpu.py -x -q 2>&1 | tail -3 | cut -c1-300
timeout 300 python tools/time_vit.py 2>&1 | tail -1 | tee gpurun_out/time_vit4.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench13.json; tail -2 gpurun_out/bench13.err | cut -c1-200
