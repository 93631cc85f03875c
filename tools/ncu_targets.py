"""One launch of each hot kernel between cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full`."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from self_corr_pose_b200 import synthetic
from self_corr_pose_b200.hotpath import HotPath, default_opts
from self_corr_pose_b200.model.module.renderer import Renderer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
opts = default_opts(batch_size=B // 4, repeat=4)
v, f = synthetic.uv_sphere()
hot = HotPath(opts, torch.from_numpy(v), torch.from_numpy(f), device='cuda', overlap_vit=False)
data, enc = synthetic.make_batch(opts, v, f, B, device='cuda', seed=0, renderer=Renderer(opts, hot.mesh))
for _ in range(2):
    hot.step(data, enc)
torch.cuda.synchronize()
torch.cuda.profiler.start()
hot.step(data, enc)          # the real step: every kernel with its real inputs (bench.py's workload, eager launch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
