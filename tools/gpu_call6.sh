#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench4.json; tail -3 gpurun_out/bench4.err
