"""One real training step between cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full`: every kernel
with its real inputs (bench.py's default workload, eager launch so that every kernel is a separate profiled launch)."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_corr_pose_b200 import synthetic
from self_corr_pose_b200.hotpath import default_opts
from self_corr_pose_b200.model.trainer import Trainer
from self_corr_pose_b200.model.module.renderer import Renderer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.backends.cudnn.benchmark = True
opts = default_opts(batch_size=B // 4, repeat=4, shape_prior_path='synthetic:uv1280')
tr = Trainer(opts)
model = tr.define_model()
model.overlap_vit = model.overlap_rotation = False            # serialised kernels: one stream
v, f = synthetic.uv_sphere()
batch = synthetic.make_trainer_batch(opts, v, f, B, device=tr.device, seed=0, renderer=Renderer(opts, model.mesh))
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
