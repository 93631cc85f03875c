#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_losses_gpu.py tests/test_hotpath_gpu.py tests/test_model_gpu.py -x -q -s 2>&1 | grep -E "PARITY cycle|PARITY hotpath|passed|failed|Error|error|assert" | cut -c1-400 | head -40 | tee gpurun_out/pytest_cycle.log
timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench8.json; tail -3 gpurun_out/bench8.err | cut -c1-300
