"""Synthetic inputs for tests and bench.py (no dataset or checkpoint is available offline).

Meshes: the category priors of the reference (config/<cat>_wild6d/<cat>.obj, shipped as the
package asset self_corr_pose_b200/data/prior_meshes.npz), a UV sphere with exactly 1280 vertices / 2556 faces
(the "1280-vertex category mesh" of BASELINE.json) and icospheres.
Scenes: SURVEY.md section 8(d) configs 0-2.
"""
import math
import os

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRIOR_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'prior_meshes.npz')


def uv_sphere(rings=18, segments=71):
    """Closed genus-0 sphere: rings*segments + 2 vertices, 2*rings*segments faces (1280 / 2556)."""
    v = [[0.0, 1.0, 0.0]]
    for r in range(1, rings + 1):
        th = math.pi * r / (rings + 1)
        for s in range(segments):
            ph = 2 * math.pi * s / segments
            v.append([math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)])
    v.append([0.0, -1.0, 0.0])
    f = []
    ring = lambda r, s: 1 + r * segments + (s % segments)
    south = len(v) - 1
    for s in range(segments):
        f.append([0, ring(0, s + 1), ring(0, s)])
        f.append([south, ring(rings - 1, s), ring(rings - 1, s + 1)])
    for r in range(rings - 1):
        for s in range(segments):
            a, b, c, d = ring(r, s), ring(r, s + 1), ring(r + 1, s), ring(r + 1, s + 1)
            f.append([a, b, d])
            f.append([a, d, c])
    return np.asarray(v, np.float32), np.asarray(f, np.int64)


def icosphere(subdivisions=3):
    """Unit icosphere: 10*4^s + 2 vertices, 20*4^s faces (642 / 1280 for s=3)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]]
    f = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
         [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
         [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    v = [list(np.asarray(p, np.float64) / np.linalg.norm(p)) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (np.asarray(v[a]) + np.asarray(v[b])) / 2
                v.append(list(m / np.linalg.norm(m)))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = nf
    return np.asarray(v, np.float32), np.asarray(f, np.int64)


def load_prior(category='laptop', normalise=True):
    """Category prior mesh; normalisation of model/module/mesh.py:70-71 (centre, divide by max |coord|)."""
    z = np.load(PRIOR_FIXTURE)
    v, f = z[category + '_v'].astype(np.float32), z[category + '_f'].astype(np.int64)
    if normalise:
        v = v - v.mean(0, keepdims=True)
        v = v / np.abs(v).max()
    return v, f


def rot_xyz(ax=0.0, ay=0.0, az=0.0):
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return (rx @ ry @ rz).astype(np.float32)


def random_poses(B, generator, device='cpu', tz=5.0):
    """Config-2 style poses: random rotations (Gram-Schmidt of N(0,1)), t_z = 5 +- 0.5, t_xy +- 0.05."""
    a = torch.randn(B, 3, 3, generator=generator)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r, dim1=1, dim2=2))[:, None, :]
    q[:, :, 2] *= torch.det(q)[:, None]
    t = torch.zeros(B, 1, 3)
    t[:, 0, 2] = tz + (torch.rand(B, generator=generator) - 0.5)
    t[:, 0, :2] = (torch.rand(B, 2, generator=generator) - 0.5) * 0.1
    return q.to(device), t.to(device)


def default_intrinsics(B, device='cpu'):
    """fp64 NDC intrinsics: focal 3.7, principal point 0."""
    foc = torch.full((B, 2), 3.7, dtype=torch.float64, device=device)
    pp = torch.zeros(B, 2, dtype=torch.float64, device=device)
    return foc, pp


def make_batch(opts, verts, faces, B, device='cuda', seed=0, renderer=None):
    """Config-2 style synthetic scene batch: per image a random pose; `mask` = hard
    silhouette and `depth` = rendered depth x 1000 (0 outside) of the prior mesh at a slightly perturbed pose,
    `img` = vertex-colour render + noise; random unit-norm image / mesh features stand in for the encoder.
    `renderer`: a model.module.renderer.Renderer used (untimed) to rasterise the ground truth.
    Returns (data, enc): data = (img, mask, depth, foc_crop, pp_crop), enc = leaf tensors requiring grad."""
    g = torch.Generator().manual_seed(seed)
    N = verts.shape[0]
    v = torch.from_numpy(verts) if not torch.is_tensor(verts) else verts
    f = torch.from_numpy(faces) if not torch.is_tensor(faces) else faces
    rot, trans = random_poses(B, g)
    foc, pp = default_intrinsics(B)
    pred_v = v[None] + 0.01 * torch.randn(B, N, 3, generator=g)
    colors = torch.rand(B, N, 3, generator=g)
    # ground-truth pose: a small perturbation of the predicted one
    d_rot, _ = random_poses(B, g)
    rot_gt = rot.bmm(torch.matrix_exp(0.05 * (d_rot - d_rot.transpose(1, 2))))
    trans_gt = trans + 0.02 * torch.randn(B, 1, 3, generator=g)
    C, P = opts.n_corr_feat, opts.corr_h * opts.corr_w
    img_feat = torch.nn.functional.normalize(torch.randn(B, C, P, generator=g), 2, 1)
    mesh_feat = torch.nn.functional.normalize(torch.relu(torch.randn(B, N, C, generator=g)), 2, -1)
    noise = 0.05 * torch.rand(B, 3, opts.img_size, opts.img_size, generator=g)

    to = lambda t: t.to(device)
    with torch.no_grad():
        fb = to(f)[None].repeat(B, 1, 1)
        gt_v = to(v)[None].repeat(B, 1, 1)
        (mask_r, tex_r, depth_r, _, _, _, depth_mask, _, _) = renderer.render_all(
            gt_v, fb, to(colors), to(foc), to(pp), to(rot_gt), to(trans_gt))
        mask = (mask_r > 0.5).float()
        depth = depth_r * 1000.0 * mask
        img = (tex_r + to(noise)).clamp(0, 1)
    data = (img.contiguous(), mask.contiguous(), depth.contiguous(), to(foc), to(pp))
    enc = tuple(to(t).clone().requires_grad_(True) for t in (img_feat, mesh_feat, pred_v, rot, trans))
    return data, enc


def make_trainer_batch(opts, verts, faces, B, device='cuda', seed=0, renderer=None):
    """The dict a Wild6D dataloader batch would hold (data/dataset_wild6d.py:165-182), synthetic: pixel-unit crop
    intrinsics in fp64 (focal 3.7*128, principal point 128), mask/depth with a channel axis."""
    data, _ = make_batch(opts, verts, faces, B, device=device, seed=seed, renderer=renderer)
    img, mask, depth, foc, pp = data
    half = opts.img_size / 2.
    return {'img': img, 'mask': mask[:, None], 'depth': depth[:, None],
            'center': torch.zeros(B, 2, dtype=torch.int64), 'length': torch.full((B, 2), opts.img_size, dtype=torch.int64),
            'foc': (foc * half).double(), 'pp': ((pp + 1) * half).double(),
            'foc_crop': (foc * half).double(), 'pp_crop': ((pp + 1) * half).double(),
            'idx': torch.arange(B)[:, None]}
