"""Generates tests/golden/matchconf_golden.npz by running the REFERENCE's `Correspondence.match` in EVALUATION mode
(model/module/correspondence.py:36-73, the forward-backward consistency confidence of :57-69) on the CPU in the build
container, with the module's unavailable imports stubbed out and `.cuda()` mapped to the identity.
Run:  python tests/golden/make_matchconf_golden.py      (needs /root/reference; the .npz is committed)
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'third-party'))
for name in ['soft_renderer', 'pytorch3d', 'pytorch3d.structures', 'pytorch3d.loss', 'pytorch3d.ops',
             'pytorch3d.ops.knn', 'pytorch3d.structures.pointclouds']:
    sys.modules[name] = types.ModuleType(name)
sys.modules['pytorch3d.ops.knn'].knn_gather = sys.modules['pytorch3d.ops.knn'].knn_points = None
sys.modules['pytorch3d.structures.pointclouds'].Pointclouds = None
torch.Tensor.cuda = lambda self, *a, **k: self

from model.module.correspondence import Correspondence  # noqa: E402

out = {}
g = torch.Generator().manual_seed(13)
B, hf, wf, N, C, H = 3, 16, 16, 90, 64, 64
opts = types.SimpleNamespace(tau_img=10., tau_mesh=10., topk_img=100, topk_mesh=100, corr_h=hf, corr_w=wf, train=False,
                             n_corr_feat=C, img_size=H)
corr = Correspondence(opts)
# features with real structure (pixel features = noisy copies of vertex features) so that the confidence map is not flat
mesh_feat = torch.nn.functional.normalize(torch.relu(torch.randn(B, N, C, generator=g)), 2, -1)
pick = torch.randint(0, N, (B, hf * wf), generator=g)
img_feat = mesh_feat.gather(1, pick[:, :, None].expand(-1, -1, C)).permute(0, 2, 1)
img_feat = torch.nn.functional.normalize(img_feat + 0.3 * torch.randn(B, C, hf * wf, generator=g), 2, 1)
yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, H), indexing='ij')
mask = ((xx ** 2 + yy ** 2) < 0.7 ** 2).float()[None].repeat(B, 1, 1)
mask[1] = ((xx - 0.2) ** 2 + yy ** 2 < 0.5 ** 2).float()
pred_v = torch.randn(B, N, 3, generator=g)
pointcorr, match, imatch, conf = corr.match(img_feat, mesh_feat, mask, pred_v)
out.update(img_feat=img_feat, mesh_feat=mesh_feat, mask=mask, pred_v=pred_v, match=match, imatch=imatch, match_conf=conf)
path = os.path.join(ROOT, 'tests', 'golden', 'matchconf_golden.npz')
np.savez_compressed(path, **{k: v.detach().numpy() for k, v in out.items()})
print('wrote', path, os.path.getsize(path), 'bytes; conf: zero fraction %.3f, max %.3f' % (float((conf == 0).float().mean()), float(conf.max())))
