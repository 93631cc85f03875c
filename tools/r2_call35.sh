#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== corr tcgen05 forward (A-stationary, compacted): parity"
  timeout 240 python -m pytest tests/test_corr_gpu.py -m gpu -q -s -x -k "tc_forward or rotation or golden" 2>&1 | grep -E "PARITY|passed|failed|Error|error" | cut -c1-400 | tail -14
  echo "== corr forward timing"
  timeout 150 python tools/time_corr.py 2>&1 | tail -8
} 2>&1 | tee gpurun_out/r2_call35.log
M="gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_op_read.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active"
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/corr_ncu3.csv python tools/ncu_corr.py > gpurun_out/corr_ncu3.log 2>&1
echo "ncu rc=$?"
