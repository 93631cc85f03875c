#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== full GPU suite"
  timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  echo "== corr timing"
  timeout 150 python tools/time_corr.py 2>&1 | tail -4
  echo "== bench (no cpu baseline, no kernel breakdown)"
  timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d.get('hotpath', {}).get('ms_per_step'))"
  echo "== bench legacy corr forward"
  SCP_CORR_FWD=legacy timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ('value','ms_per_step','clocks')}, d['e2e']['value'], d.get('hotpath', {}).get('ms_per_step'))"
} 2>&1 | tee gpurun_out/r2_call36.log
