"""Times the batched inference pose fit (self_corr_pose_b200/model/pose_fit.py) on cuda:0 against
the reference's one-image / one-round-at-a-time formulation (oracle/posefit.py) on the host cores, and checks the two
against each other on the same batch.   python tools/time_posefit.py [--batch 32] [--size 256] [--oracle-images 4]
Wall-clock per batch (the fit contains host synchronisations by construction), after one warm-up call."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import argparse
import os
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import posefit as O  # noqa: E402
from self_corr_pose_b200.model.pose_fit import PoseFitter  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--size', type=int, default=256)
ap.add_argument('--verts', type=int, default=1280)
ap.add_argument('--oracle-images', type=int, default=4)
ap.add_argument('--repeat', type=int, default=5)
ap.add_argument('--device', default='cuda')
a = ap.parse_args()

B, size, N = a.batch, a.size, a.verts
g = torch.Generator().manual_seed(0)
foc = (3.5 + 0.3 * torch.rand(B, 2, generator=g)).double()
pp = (0.05 * torch.randn(B, 2, generator=g)).double()
match = torch.rand(B, 3, size, size, generator=g) - 0.5
A = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
cam = 300 * torch.einsum('bchw,bcd->bdhw', match, A) + torch.tensor([0., 0., 900.])[None, :, None, None]
depth = (cam[:, 2] + 3 * torch.randn(B, size, size, generator=g)) * (torch.rand(B, size, size, generator=g) > 0.1)
yy, xx = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing='ij')
mask = ((xx ** 2 + yy ** 2) < 0.6 ** 2).float()[None].repeat(B, 1, 1)          # ~28 % of the crop is foreground
conf = torch.rand(B, 1, size, size, generator=g)
conf[conf < 0.4] = 0
pred_v = torch.rand(B, N, 3, generator=g) - 0.5
opts = SimpleNamespace(img_size=size)

dev = a.device
to = lambda t: t.to(dev)
batch = (None, to(mask), to(depth), None, None, None, None, to(foc), None, to(pp), None, None)
pred = (to(pred_v), None, None, None, to(match), to(conf))
fitter = PoseFitter(opts, device=dev)
torch.manual_seed(1)
got = fitter.pose_fitting(batch, pred)
if dev != 'cpu':
    torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.repeat):
    fitter.pose_fitting(batch, pred)
if dev != 'cpu':
    torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / a.repeat

k = min(a.oracle_images, B)
torch.manual_seed(1)
t0 = time.perf_counter()
want = O.pose_fitting(mask[:k], depth[:k], match[:k], conf[:k], foc[:k], pp[:k], pred_v[:k], torch.eye(3).reshape(-1), size)
t_cpu = (time.perf_counter() - t0) / k
err = max(float((x[:k].cpu() - y).abs().max() / y.abs().max()) for x, y in zip(got, want))
n = int(((depth > 0)[:, None] * mask[:, None] * conf > 0).sum()) // B
print('pose fit: %d images of %dx%d, ~%d correspondences each' % (B, size, size, n))
print('  batched, ' + dev + '            %8.2f ms / batch   %8.3f ms / image' % (1e3 * t_gpu, 1e3 * t_gpu / B))
print('  reference formulation, host %6.2f ms / image  (%d images, %d threads)' % (1e3 * t_cpu, k, torch.get_num_threads()))
print('  max relative difference on the first %d images: %.2e' % (k, err))
