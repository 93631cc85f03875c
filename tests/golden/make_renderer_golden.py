"""Generates tests/golden/renderer_golden.npz by running the REFERENCE's own model/module/renderer.py
(Renderer.render_all: the four renders through loss_utils.render + the reference soft_renderer front-end, imatch_gt,
depth_weight) on the CPU in the build container.  The CUDA operator is replaced by a deterministic stand-in
(`fake_rasterize`, defined identically in tests/test_oracle_renderer.py) whose output depends on every argument it is
given, so the nine outputs of render_all pin the whole host-side chain: camera transform, fp64 pinhole projection, y flip,
depth texture, look_at, face gathers, renderer settings and background colours, the order of the four operator calls, the
output slicing, grid_sample and the visibility weight.
Run:  python tests/golden/make_renderer_golden.py      (needs /root/reference; the .npz is committed)
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'third-party', 'softras'))
for name in ['soft_renderer.cuda', 'soft_renderer.cuda.soft_rasterize', 'soft_renderer.cuda.load_textures',
             'soft_renderer.cuda.create_texture_image', 'soft_renderer.cuda.voxelization', 'skimage', 'skimage.io',
             'pytorch3d', 'pytorch3d.structures', 'pytorch3d.loss', 'pytorch3d.ops', 'pytorch3d.ops.knn',
             'pytorch3d.structures.pointclouds']:
    sys.modules[name] = types.ModuleType(name)
sys.modules['skimage.io'].imread = sys.modules['skimage.io'].imsave = None
sys.modules['pytorch3d.ops.knn'].knn_gather = sys.modules['pytorch3d.ops.knn'].knn_points = None
sys.modules['pytorch3d.structures.pointclouds'].Pointclouds = None
torch.Tensor.cuda = lambda self, *a, **k: self

import soft_renderer.rasterizer as ref_rasterizer           # noqa: E402
from model.module.renderer import Renderer                  # noqa: E402  (the reference class)

CALLS = []


def fake_rasterize(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
                   eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb='softmax',
                   aggr_func_alpha='prod', texture_type='surface'):
    """deterministic stand-in for the SoftRas operator: smooth images that depend on every argument"""
    B = face_vertices.shape[0]
    CALLS.append([image_size, list(background_color), near, far, bool(fill_back), eps, sigma_val, dist_func, dist_eps,
                  gamma_val, aggr_func_rgb, aggr_func_alpha, texture_type])
    s = face_vertices.reshape(B, -1).sum(1) * 0.37 + textures.reshape(B, -1).sum(1) * 0.11 + sigma_val * 1e3 + \
        gamma_val * 1e2 + (0.5 if aggr_func_rgb == 'hard' else 0.) + sum(background_color) * 0.25
    ys, xs = torch.meshgrid(torch.arange(image_size, dtype=torch.float32), torch.arange(image_size, dtype=torch.float32),
                            indexing='ij')
    c = torch.arange(4, dtype=torch.float32)[None, :, None, None]
    img = torch.sin(0.1 * (c + 1) * xs[None, None] + 0.07 * ys[None, None] + s[:, None, None, None]) * 0.5 + 0.5
    img = img * torch.tensor([1., 1., 6., 1.])[None, :, None, None]          # channel 2 plays the depth
    return img


ref_rasterizer.srf.soft_rasterize = fake_rasterize

from self_corr_pose_b200 import synthetic                    # noqa: E402

g = torch.Generator().manual_seed(13)
v, f = synthetic.icosphere(1)
B, N = 4, v.shape[0]      # not 3: the reference's look_at calls torch.cross without dim, which picks the batch axis when B == 3
opts = types.SimpleNamespace(img_size=32, use_depth=True)
mesh = types.SimpleNamespace(mean_v=torch.from_numpy(v), faces=torch.from_numpy(f), texture_type='vertex')
pred_v = (torch.from_numpy(v)[None] + 0.02 * torch.randn(B, N, 3, generator=g)).requires_grad_(True)
faces = torch.from_numpy(f)[None].repeat(B, 1, 1)
tex = torch.rand(B, N, 3, generator=g)
rot, trans = synthetic.random_poses(B, g)
foc = (3.7 + 0.3 * torch.rand(B, 2, generator=g)).double()
pp = (0.1 * (torch.rand(B, 2, generator=g) - 0.5)).double()

outs = Renderer(opts, mesh).render_all(pred_v, faces, tex, foc, pp, rot, trans, None)
names = ('mask_render', 'tex_render', 'depth_render', 'match_gt', 'imatch_gt', 'tex_mask', 'depth_mask', 'match_mask',
         'depth_weight')
out = dict(pred_v=pred_v.detach(), faces=faces, tex=tex, rot=rot, trans=trans, foc=foc, pp=pp, v=torch.from_numpy(v),
           f=torch.from_numpy(f), calls=np.array(repr(CALLS)))
for n, t in zip(names, outs):
    out['o_' + n] = t.detach()
    print(n, tuple(t.shape))
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'renderer_golden.npz'),
                    **{k: (x.detach().numpy() if torch.is_tensor(x) else np.asarray(x)) for k, x in out.items()})
