"""Generates tests/golden/loss_golden.npz by running the REFERENCE's own loss / camera functions
(model/util/loss_utils.py: compute_mask_loss, compute_texture_loss, compute_depth_loss, compute_match_loss,
compute_imatch_loss, pinhole_cam, LaplacianLoss, divide_by_*) on the CPU in the build container, values AND gradients
(autograd of the reference statements), with the module's unavailable imports (soft_renderer, pytorch3d) stubbed out --
none of the functions exercised here touches them.
Run:  python tests/golden/make_loss_golden.py      (needs /root/reference; the .npz is committed)
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'third-party'))
for name in ['soft_renderer', 'pytorch3d', 'pytorch3d.structures', 'pytorch3d.loss', 'pytorch3d.ops',
             'pytorch3d.ops.knn', 'pytorch3d.structures.pointclouds']:
    sys.modules[name] = types.ModuleType(name)
sys.modules['pytorch3d.ops.knn'].knn_gather = sys.modules['pytorch3d.ops.knn'].knn_points = None
sys.modules['pytorch3d.structures.pointclouds'].Pointclouds = None

from model.util import loss_utils as R  # noqa: E402  (the reference module)

g = torch.Generator().manual_seed(21)
rand = lambda *s: torch.rand(*s, generator=g)
randn = lambda *s: torch.randn(*s, generator=g)
B, H, hf, N = 3, 32, 8, 50
out = {}

# renders as the SoftRas operator returns them: (B,4,H,H)
r_depth = torch.cat([rand(B, 2, H, H), 3 * rand(B, 1, H, H), (rand(B, 1, H, H) - 0.3).clamp(0, 1)], 1)
r_tex = rand(B, 4, H, H)
r_nocs = torch.cat([rand(B, 3, H, H) - 0.5, (rand(B, 1, H, H) - 0.4).clamp(0, 1)], 1)
match_lr = rand(B, hf * hf, 3) - 0.5
img = rand(B, 3, H, H)
mask = (rand(B, H, H) > 0.4).float()
depth = 3 * rand(B, H, H) * (rand(B, H, H) > 0.2).float()
w = rand(B, 4) + 0.5
out.update(r_depth=r_depth, r_tex=r_tex, r_nocs=r_nocs, match_lr=match_lr, img=img, mask=mask, depth=depth, w=w)

leaves = [t.clone().requires_grad_(True) for t in (r_depth, r_tex, match_lr)]
rd, rt, ml = leaves
match = F.interpolate(ml.reshape(B, hf, hf, 3).permute(0, 3, 1, 2), (H, H), mode='nearest')   # correspondence.py:71
l_mask = R.compute_mask_loss(img, mask, rd[:, 3])
l_tex = R.compute_texture_loss(img, mask, rt[:, :3], rt[:, 3])
l_depth, depth_diff = R.compute_depth_loss(depth.clone(), rd[:, 2].clone(), rd[:, 3], mask)
l_match = R.compute_match_loss(match, r_nocs[:, :3], r_nocs[:, 3], mask)
losses = torch.stack([l_mask, l_tex, l_depth, l_match], 1)
(losses * w).sum().backward()
out.update(losses=losses.detach(), g_r_depth=rd.grad, g_r_tex=rt.grad, g_match_lr=ml.grad)

# imatch loss
imatch, imatch_gt, dw = randn(B, 2, N), randn(B, 2, N), rand(B, N)
out.update(imatch=imatch, imatch_gt=imatch_gt, dw=dw, l_imatch=R.compute_imatch_loss(imatch, imatch_gt, dw))

# camera: verts.bmm(R) + t -> pinhole_cam (fp64 intrinsics) -> y flip, as in render()
verts = randn(B, N, 3) * 0.3
rot = torch.linalg.qr(randn(B, 3, 3))[0]
trans = torch.tensor([[[0.02, -0.01, 5.0]]]).repeat(B, 1, 1) + 0.05 * randn(B, 1, 3)
foc = (3.7 + 0.3 * rand(B, 2)).double()
pp = (0.1 * (rand(B, 2) - 0.5)).double()
cam_leaves = [t.clone().requires_grad_(True) for t in (verts, rot, trans)]
sv = cam_leaves[0].bmm(cam_leaves[1].clone()) + cam_leaves[2].clone()
sv = R.pinhole_cam(sv, pp, foc)
sv[:, :, 1] *= -1
w_sv = randn(B, N, 3)
(sv * w_sv).sum().backward()
out.update(c_verts=verts, c_rot=rot, c_trans=trans, c_foc=foc, c_pp=pp, c_w=w_sv, c_screen=sv.detach(),
           c_g_verts=cam_leaves[0].grad, c_g_rot=cam_leaves[1].grad, c_g_trans=cam_leaves[2].grad)

# Laplacian smoothness loss on an icosphere
from self_corr_pose_b200 import synthetic  # noqa: E402
v, f = synthetic.icosphere(1)
lap = R.LaplacianLoss(torch.from_numpy(v), torch.from_numpy(f), average=True)
x = (torch.from_numpy(v)[None] + 0.05 * randn(B, v.shape[0], 3)).requires_grad_(True)
ll = lap(x)
ll.backward()
out.update(lap_v=torch.from_numpy(v), lap_f=torch.from_numpy(f), lap_x=x.detach(), lap_loss=ll.detach(), lap_g=x.grad)

# batch pairing
t = torch.arange(24.).reshape(12, 2)
for name in ('divide_by_frame', 'divide_by_instance', 'divide_by_both'):
    s, tt = getattr(R, name)(t, 3, 4)
    out[name + '_src'], out[name + '_tgt'] = s, tt

np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'loss_golden.npz'),
                    **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
print('wrote', len(out), 'arrays')
