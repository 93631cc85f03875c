"""Generates tests/golden/posefit_golden.npz by running the REFERENCE's inference pose fit on the CPU in the build
container: `estimateSimilarityTransform` of model/util/umeyama.py (imported as is) on seeded correspondence sets, and
`Tester.pose_fitting` (model/tester.py:324-427, compiled from its source without importing the module, which pulls in
trimesh / kornia / the dataset) on a seeded synthetic batch.  The random generator the reference samples from is the
global CPU one: every case records the seed set right before the call and the first random integers drawn right after
it, which pins how many random numbers the reference consumed.
Run:  python tests/golden/make_posefit_golden.py      (needs /root/reference; the .npz is committed)
"""
import ast
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

spec = importlib.util.spec_from_file_location('ref_umeyama', os.path.join(REF, 'model/util/umeyama.py'))
ref_u = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_u)


def reference_pose_fitting():
    from self_corr_pose_b200.model.util.loss_utils import pinhole_cam      # pinned to the reference by loss_golden.npz
    tree = ast.parse(open(os.path.join(REF, 'model/tester.py')).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'Tester')
    fn = next(f for f in cls.body if isinstance(f, ast.FunctionDef) and f.name == 'pose_fitting')
    ns = {'torch': torch, 'estimateSimilarityTransform': ref_u.estimateSimilarityTransform, 'pinhole_cam': pinhole_cam}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'reference_tester', 'exec'), ns)
    return ns['pose_fitting']


def correspondences(n, seed, outlier, noise, shrink):
    """model points in a unit cube, camera points = similarity transform of them (millimetres) + noise + gross outliers;
    `shrink` scales both sets down so that the residuals fall below the stop threshold and the loop ends early"""
    g = torch.Generator().manual_seed(seed)
    src = torch.rand(n, 3, generator=g) - 0.5
    A = torch.linalg.qr(torch.randn(3, 3, generator=g))[0]
    tgt = (300 + 100 * torch.rand(1, generator=g)) * src @ A + torch.tensor([10., -20., 800.])
    tgt = tgt + noise * torch.randn(n, 3, generator=g)
    k = int(outlier * n)
    tgt[:k] += 200 * torch.randn(k, 3, generator=g)
    return (src * 1e-4, tgt * 1e-7) if shrink else (src, tgt)


out = {}
CASES = [(300, 0, 0.2, 2.0, False), (1200, 1, 0.3, 3.0, False), (400, 5, 0.0, 0.05, True), (7, 8, 0.0, 0.5, False),
         (650, 9, 0.2, 0.3, True), (5000, 10, 0.25, 2.0, False)]
out['fit_cases'] = np.array(len(CASES))
for i, c in enumerate(CASES):
    src, tgt = correspondences(*c)
    torch.manual_seed(100 + i)
    scales, rot, trans, transform = ref_u.estimateSimilarityTransform(src, tgt)
    out.update({'fit%d_src' % i: src.numpy(), 'fit%d_tgt' % i: tgt.numpy(), 'fit%d_scales' % i: scales.numpy(),
                'fit%d_rotation' % i: rot.numpy(), 'fit%d_translation' % i: trans.numpy(),
                'fit%d_transform' % i: transform.numpy(), 'fit%d_next_draw' % i: torch.randint(0, 1 << 30, (4,)).numpy()})

# ---- Tester.pose_fitting on a synthetic evaluation batch -------------------------------------------------------------
torch.Tensor.cuda = lambda self, *a, **k: self
size, B, N = 32, 6, 60
g = torch.Generator().manual_seed(0)
base_rot = [0, 1, 0, -1, 0, 0, 0, 0, 1]
foc = (3.5 + 0.3 * torch.rand(B, 2, generator=g)).double()
pp = (0.05 * torch.randn(B, 2, generator=g)).double()
match = torch.rand(B, 3, size, size, generator=g) - 0.5
A = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
cam = 300 * torch.einsum('bchw,bcd->bdhw', match, A) + torch.tensor([0., 0., 900.])[None, :, None, None]
depth = (cam[:, 2] + 3 * torch.randn(B, size, size, generator=g)) * (torch.rand(B, size, size, generator=g) > 0.1)
mask = (torch.rand(B, size, size, generator=g) > 0.4).float()
conf = torch.rand(B, 1, size, size, generator=g)
conf[conf < 0.3] = 0
mask[3] = 0                       # no confident foreground pixel: the fit raises, default pose
mask[4].reshape(-1)[3:] = 0       # at most three correspondences
pred_v = torch.rand(B, N, 3, generator=g) - 0.5
grid = torch.Tensor(np.array(np.meshgrid(range(size), range(size)))).reshape(2, -1) + 0.5     # tester.py:134-137
grid = (grid / (size / 2) - 1).reshape(-1)
ref_self = SimpleNamespace(opts=SimpleNamespace(img_size=size), meshgrid=grid,
                           base_rot=torch.tensor(base_rot, dtype=torch.float32).reshape(1, 3, 3))
batch = (torch.zeros(B, 3, size, size), mask, depth, None, None, None, None, foc, None, pp, None, None)
pred = (pred_v, None, None, None, match, conf)
torch.manual_seed(9)
bbox, verts, rotation, translation = reference_pose_fitting()(ref_self, batch, pred)
out.update(pose_size=np.array(size), pose_base_rot=np.array(base_rot, dtype=np.float32), pose_foc=foc.numpy(),
           pose_pp=pp.numpy(), pose_match=match.numpy(), pose_depth=depth.numpy(), pose_mask=mask.numpy(),
           pose_conf=conf.numpy(), pose_pred_v=pred_v.numpy(), pose_bbox=bbox.numpy(), pose_verts=verts.numpy(),
           pose_rotation=rotation.numpy(), pose_translation=translation.numpy(),
           pose_next_draw=torch.randint(0, 1 << 30, (4,)).numpy())
path = os.path.join(ROOT, 'tests', 'golden', 'posefit_golden.npz')
np.savez_compressed(path, **out)
print('wrote', path, os.path.getsize(path), 'bytes;', {k: v.shape for k, v in out.items() if k.startswith('pose_')})
