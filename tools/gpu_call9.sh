#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_geom_gpu.py tests/test_hotpath_gpu.py tests/test_losses_gpu.py tests/test_model_gpu.py -x -q -s 2>&1 | grep -E "PARITY project|PARITY fused|passed|failed|Error|error|assert" | cut -c1-300 | head -40 | tee gpurun_out/pytest_geom.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench6.json; tail -3 gpurun_out/bench6.err | cut -c1-300
