#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hotpath_gpu.py tests/test_losses_gpu.py -x -q 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown --no-overlap > gpurun_out/bench7_noov.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; head -c 250 gpurun_out/bench7_noov.json; echo
timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown > gpurun_out/bench7_ov.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; head -c 250 gpurun_out/bench7_ov.json; echo
timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown --no-graph > gpurun_out/bench7_eager.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; head -c 250 gpurun_out/bench7_eager.json; echo
