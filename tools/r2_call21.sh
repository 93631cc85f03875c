#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/sweep_ncu.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
    -k regex:"forward_kernel|backward_face_kernel|corr_fwd_kernel|corr_bwd_rows_kernel" python tools/sweep_ncu.py run > gpurun_out/sweep_ncu.log 2>&1
echo "rc=$?"; wc -l gpurun_out/sweep_ncu.csv; tail -3 gpurun_out/sweep_ncu.log
echo "== bench kernel breakdown with proper operands"
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline'])
for k in d['kernels']: print('%-70s %8.3f ms x%d  %8.1f %s frac %.4f' % (k['kernel'][:70], k['ms'], k['launches_per_step'], k['achieved'], k['unit'], k['frac']))"
