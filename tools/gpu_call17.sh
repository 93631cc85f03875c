#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench11.json; tail -2 gpurun_out/bench11.err | cut -c1-200
