"""One hot-path step under torch.profiler: kernel-time table (names + totals) for finding glue hot spots."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from self_corr_pose_b200.hotpath import HotPath, default_opts
from self_corr_pose_b200.model.module.renderer import Renderer
from self_corr_pose_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
opts = default_opts(batch_size=B // 4, repeat=4)
v, f = synthetic.uv_sphere()
hot = HotPath(opts, torch.from_numpy(v), torch.from_numpy(f), device='cuda')
data, enc = synthetic.make_batch(opts, v, f, B, device='cuda', seed=0, renderer=Renderer(opts, hot.mesh))
for _ in range(3):
    hot.step(data, enc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    hot.step(data, enc)
e1.record(); torch.cuda.synchronize()
print('step ms', e0.elapsed_time(e1) / 5)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    hot.step(data, enc)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70))
