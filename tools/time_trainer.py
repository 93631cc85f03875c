"""Times Trainer.step (full model: ResNet-18 encoder x2 passes, heads, hot path, AdamW) with CUDA events."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_corr_pose_b200 import synthetic
from self_corr_pose_b200.hotpath import default_opts
from self_corr_pose_b200.model.trainer import Trainer
from self_corr_pose_b200.model.module.renderer import Renderer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.backends.cudnn.benchmark = True
opts = default_opts(batch_size=B // 4, repeat=4)
tr = Trainer(opts)
model = tr.define_model()
v, f = synthetic.load_prior('laptop')
batch = synthetic.make_trainer_batch(opts, v, f, B, device=tr.device, seed=0, renderer=Renderer(opts, model.mesh))
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    total, aux, _ = tr.step(batch)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({'trainer_step_ms': ms, 'images_per_sec': B / ms * 1e3, 'B': B, 'loss': float(total),
                  'aux': {k: float(x) for k, x in aux.items()}}))
