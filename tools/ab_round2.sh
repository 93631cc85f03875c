#!/bin/bash
# Round-2 opener.  In the container:   tools/ab_round2.sh build     (compiles the candidate variants, ~1 min)
# then on the GPU box:                 gpurun --timeout 1500 -- 'bash tools/ab_round2.sh run'
# Candidates and what they change: DESIGN.md section 7.  Variants are opt-in (SCP_LIB_VARIANT); the product library is
# untouched.  Everything is written to gpurun_out/ab_round2.log as well.
set -u
if [ "${1:-}" = "build" ]; then
  python -m self_corr_pose_b200.build
  python -m self_corr_pose_b200.build --variant early_qk -DSCP_FA2_EARLY_QK=1
  python -m self_corr_pose_b200.build --variant fwd2px   -DSCP_SOFTRAS_FWD_2PX=1
  python -m self_corr_pose_b200.build --variant facesmem -DSCP_SOFTRAS_FACE_SMEM=1 -DSCP_SOFTRAS_FACE_CTAS=10
  python -m self_corr_pose_b200.build --variant face16x2 -DSCP_SOFTRAS_FACE_SMEM=1 -DSCP_SOFTRAS_FACE_CTAS=10 -DSCP_SOFTRAS_FACE_BW=16
  python -m self_corr_pose_b200.build --variant softras_all -DSCP_SOFTRAS_FWD_2PX=1 -DSCP_SOFTRAS_FACE_SMEM=1 -DSCP_SOFTRAS_FACE_CTAS=10 -DSCP_SOFTRAS_FACE_BW=16
  ls -la self_corr_pose_b200/*.so
  exit 0
fi
mkdir -p gpurun_out
{
  echo "== pose fit (first GPU run)"
  timeout 300 python -m pytest tests/test_posefit.py -m gpu -q -rxX 2>&1 | tail -4
  timeout 300 python tools/time_posefit.py 2>&1 | tail -5
  echo "== attention: early S issue"
  timeout 900 python tools/ab_variants.py --only early_qk --tests tests/test_vit_gpu.py --time "tools/time_vit.py 64"
  echo "== SoftRas candidates (tests = SoftRas parity suite; timing = whole-step bench line incl. the kernel breakdown)"
  timeout 1500 python tools/ab_variants.py --only fwd2px,facesmem,face16x2,softras_all --tests tests/test_softras_gpu.py \
      --time "bench.py --steps 10 --warmup 3 --no-cpu-baseline"
} 2>&1 | tee gpurun_out/ab_round2.log
