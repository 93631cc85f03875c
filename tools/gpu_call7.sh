#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_corr_gpu.py tests/test_hotpath_gpu.py tests/test_losses_gpu.py tests/test_model_gpu.py -x -q -s 2>&1 | grep -E "PARITY corr|PARITY rot|passed|failed|Error|error|assert" | cut -c1-600 | head -40 | tee gpurun_out/pytest_corr.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err | cut -c1-300
