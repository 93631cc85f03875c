"""Where does Trainer.step go?  (1) wall/device time of the step, (2) torch.profiler kernel table (device time by kernel
name), (3) device time of the step's sections measured with CUDA events by monkey-patching the model's sub-calls."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import sys, os, json, time, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_corr_pose_b200 import synthetic
from self_corr_pose_b200.hotpath import default_opts
from self_corr_pose_b200.model.trainer import Trainer
from self_corr_pose_b200.model.module.renderer import Renderer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.backends.cudnn.benchmark = True
opts = default_opts(batch_size=B // 4, repeat=4)
tr = Trainer(opts)
model = tr.define_model()
v, f = synthetic.load_prior('laptop')
batch = synthetic.make_trainer_batch(opts, v, f, B, device=tr.device, seed=0, renderer=Renderer(opts, model.mesh))
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
n = 5
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    tr.step(batch)
e1.record(); torch.cuda.synchronize()
print(json.dumps({'step_ms_device': e0.elapsed_time(e1) / n, 'step_ms_wall': (time.time() - t0) / n * 1e3}))

# sections (device time between events; includes host gaps when launch-bound)
sec = collections.OrderedDict()
def wrap(obj, name, label):
    fn = getattr(obj, name)
    def w(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(*a, **k); e.record()
        sec.setdefault(label, []).append((s, e))
        return r
    setattr(obj, name, w)
wrap(model.encoder, 'forward', 'encoder.forward (ResNet+heads)')
wrap(model.corr_net, 'match_lowres', 'corr.match_lowres')
wrap(model.renderer, 'render_all_raw', 'renderer.render_all_raw')
wrap(model.mesh, 'compute_symmetry_loss', 'mesh.compute_symmetry_loss')
wrap(model.pretrain_corr_net, 'compute_cycle_loss', 'pretrain.compute_cycle_loss (incl. ViT)')
wrap(model.corr_net, 'compute_rotation_cycle_loss', 'corr.rotation_cycle (2nd encoder pass)')
wrap(tr.optim, 'step', 'optim.step')
wrap(tr, 'collect_grad', 'collect_grad')
wrap(tr.optim, 'zero_grad', 'zero_grad')
orig_model_call = tr.model.forward
def fwd(*a, **k):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = orig_model_call(*a, **k); e.record(); sec.setdefault('model.forward total', []).append((s, e)); return r
tr.model.forward = fwd
s_all = []
for _ in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); tr.step(batch); e.record(); s_all.append((s, e))
torch.cuda.synchronize()
print('section device-time (ms, mean of 3):')
for k, evs in sec.items():
    print('  %-45s %8.3f' % (k, sum(a.elapsed_time(b) for a, b in evs) / 3))
print('  %-45s %8.3f' % ('whole step', sum(a.elapsed_time(b) for a, b in s_all) / 3))

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(2):
        tr.step(batch)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = [(k.key, getattr(k, 'device_time_total', getattr(k, 'cuda_time_total', 0)) / 2e3, k.count // 2) for k in ka
        if getattr(k, 'device_type', None) is not None and str(k.device_type).endswith('CUDA')]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print('kernel device time per step: %.3f ms over %d launches' % (tot, sum(r[2] for r in rows)))
for name, ms, cnt in rows[:60]:
    print('  %8.3f ms  x%-4d %s' % (ms, cnt, name[:110]))

# by ATen operator and input shape (self device time): which host statements the long tail of small kernels comes from
ops = prof.key_averages(group_by_input_shape=True)
orow = [(o.key, str(o.input_shapes)[:70], getattr(o, 'self_device_time_total', 0) / 2e3, o.count // 2) for o in ops]
orow = [r for r in orow if r[2] > 0.15]
orow.sort(key=lambda r: -r[2])
print('ATen ops by self device time per step (> 0.15 ms):')
for name, shp, ms, cnt in orow[:70]:
    print('  %8.3f ms  x%-4d %-40s %s' % (ms, cnt, name[:40], shp))
