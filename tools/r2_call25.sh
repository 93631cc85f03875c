#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa3_fwd_kernel --launch-skip 4 -c 1 -o /tmp/fa3 python tools/time_vit.py 64 > gpurun_out/fa3_ncu.log 2>&1
echo "rc=$?"
ncu -i /tmp/fa3.ncu-rep --page raw --csv > gpurun_out/fa3_raw.csv 2>/dev/null
ncu -i /tmp/fa3.ncu-rep --page source --csv > gpurun_out/fa3_source.csv 2>/dev/null
gzip -f gpurun_out/fa3_source.csv
ls -la gpurun_out/fa3*
