#!/bin/bash
# ncu evidence of a round (run under gpurun, 1 GPU): launch list of the bench command + `--set full` of one real step.
# Only CSV exports come back (the .ncu-rep files stay in /tmp on the box: gpurun_out/ is capped at 64 MiB).
# ncu's -k filter matches the kernel BASE name (no namespace / template arguments), hence the explicit list.
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
gzip -f gpurun_out/launches.csv
KERNELS="forward_kernel|backward_kernel|backward_face_kernel|pack_kernel|corr_|gemm_bf16_tn_kernel|fa3_fwd_kernel|layernorm"
KERNELS="$KERNELS|image_loss_kernel|depth_sums|project_faces|spmm3|cycle_rows|im2col|nn_fwd_kernel|nn_bwd_kernel|jitter_norm_kernel|gray_sum_kernel"
KERNELS="$KERNELS|maxpool_fwd_kernel|maxpool_bwd_kernel|upsample_fwd_kernel|upsample2x_bwd_kernel|l2norm_fwd_kernel|l2norm_bwd_kernel"
timeout 2400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$KERNELS" \
    -o /tmp/step python tools/ncu_targets.py > gpurun_out/ncu_step.log 2>&1
echo "full rc=$?"
ncu -i /tmp/step.ncu-rep --page raw --csv > gpurun_out/step_raw.csv 2>/dev/null
for k in fa3_fwd_kernel backward_face_kernel forward_kernel; do
    ncu -i /tmp/step.ncu-rep --page source --csv --kernel-name $k --launch-count 1 > gpurun_out/${k}_source.csv 2>/dev/null
    gzip -f gpurun_out/${k}_source.csv
done
gzip -f gpurun_out/step_raw.csv
ls -la gpurun_out/ | tail -8
# then, back in the build container:
#   python tools/summarize_launches.py gpurun_out/launches.csv.gz profiles/rN_bench_launch_summary.md "<note>"
#   python tools/summarize_ncu.py gpurun_out/step_raw.csv profiles/rN_hot_kernels_ncu_full.md profiles/traffic.json "<note>"
