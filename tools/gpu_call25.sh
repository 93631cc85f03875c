#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_softras_gpu.py tests/test_hotpath_gpu.py tests/test_losses_gpu.py -x -q -s 2>&1 | grep -E "PARITY face|passed|failed|Error|error|assert" | cut -c1-200 | head -20
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench14.json; tail -2 gpurun_out/bench14.err | cut -c1-200
