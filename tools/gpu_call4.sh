#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_losses_gpu.py tests/test_hotpath_gpu.py tests/test_model_gpu.py -x -q -s 2>&1 | grep -E "PARITY|passed|failed|Error|error|assert" | head -60 | tee gpurun_out/pytest_losses.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench3.json; tail -3 gpurun_out/bench3.err
