#!/bin/bash
# Round-2 GPU call 5: parity at the benchmarked shape + Trainer.step parity, full GPU suite, the new bench.py arms.
set -u
mkdir -p gpurun_out
{
  echo "== new parity tests"
  timeout 1500 python -m pytest tests/test_jitter_gpu.py tests/test_hotpath_gpu.py tests/test_model_gpu.py -m gpu -q -s 2>&1 | grep -e PARITY -e passed -e failed -e Error -e "^E " | tail -60
  echo "== full GPU suite"
  timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15
  echo "== bench (trainer)"
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_trainer.json 2> gpurun_out/r2_bench_trainer.err; tail -c 6000 gpurun_out/r2_bench_trainer.json; tail -5 gpurun_out/r2_bench_trainer.err
  echo "== bench (config1)"
  timeout 600 python bench.py --workload config1 --steps 10 --warmup 3 > gpurun_out/r2_bench_config1.json 2> gpurun_out/r2_bench_config1.err; cat gpurun_out/r2_bench_config1.json; tail -5 gpurun_out/r2_bench_config1.err
  echo "== Trainer.step profile"
  timeout 600 python tools/profile_trainer.py 64 2>&1 | grep -v Warn | head -75
} 2>&1 | tee gpurun_out/r2_call5.log
