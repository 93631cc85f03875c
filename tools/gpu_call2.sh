#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_vit.log
for v in 1 2; do SCP_VIT_ATTENTION=$v timeout 300 python tools/time_vit.py 2>&1 | tail -1 | tee -a gpurun_out/time_vit.log; done
timeout 600 python tools/sweep_kernels.py gpurun_out/sweep.md > gpurun_out/sweep.log 2>&1; tail -3 gpurun_out/sweep.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err
