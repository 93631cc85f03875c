#!/bin/bash
set -u
mkdir -p gpurun_out
{
  for c4 in 8 0; do
    echo "== bench SCP_STEM_C4=$c4"
    SCP_STEM_C4=$c4 timeout 900 python bench.py --no-cpu-baseline --no-kernel-breakdown 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ('value','ms_per_step','clocks')}, d['e2e']['value'], d['loss'])"
  done
} 2>&1 | tee gpurun_out/r2_call31.log
