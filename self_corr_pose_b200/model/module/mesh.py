"""Canonical category mesh: learnable mean shape, fixed faces, vertex-colour texture lookup, symmetry loss.
API of the reference's model/module/mesh.py:29-131 (`CanonicalMesh(opts)`: `mean_v`, `faces`, `symm_rots`,
`get_texture`, `compute_symmetry_loss`).  trimesh / pytorch3d are not available offline: OBJ priors are read with a
plain parser (the shipped priors are duplicate-free, so `trimesh.load_mesh(process=True)` returns the same arrays)
and the symmetry regulariser's point sampling + 1-NN search are restated in torch: the reference delegates them to
pytorch3d 0.6.1, absent here -- the symmetry rotations and the one-way chamfer reduction are
pinned to the reference's own model/util/{symmetry,chamfer}.py with a brute-force stand-in for pytorch3d's knn_points
(tests/test_reference_host_cpu.py); the random surface sampling stays unpinned."""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _flags, synthetic


def read_obj(path):
    v, f = [], []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == 'v':
                v.append([float(x) for x in t[1:4]])
            elif t[0] == 'f':
                f.append([int(x.split('/')[0]) - 1 for x in t[1:4]])
    return np.asarray(v, np.float32), np.asarray(f, np.int64)


def get_symm_rots(division):
    rots = torch.zeros(division, 3, 3)
    for i in range(division):
        th = 2 * np.pi / division * i
        rots[i] = torch.tensor([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    return rots


def sample_faces_and_weights(verts, faces, num_samples):
    """The random part of pytorch3d.ops.sample_points_from_meshes: area-weighted face draws (multinomial with replacement)
    and uniform barycentric weights (w0, w1, w2) = (1 - sqrt(u), sqrt(u) (1 - v), sqrt(u) v).  verts (B,N,3),
    faces (B,nf,3) -> face_idx (B,S) int64, w (B,S,3).  Consumes the device generator: two rand calls."""
    B = verts.shape[0]
    idx = faces.long()
    v0, v1, v2 = (torch.gather(verts, 1, idx[:, :, k:k + 1].expand(-1, -1, 3)) for k in range(3))
    areas = 0.5 * torch.cross(v1 - v0, v2 - v0, dim=-1).norm(dim=-1)
    # area-weighted draws with replacement by inverting the cumulative areas: the distribution of
    # `areas.multinomial(num_samples, replacement=True)` (pytorch3d) without torch.multinomial's host-synchronising input
    # checks, which cannot be captured into a CUDA graph
    cdf = areas.detach().clamp_min(1e-12).cumsum(1)
    r = torch.rand(B, num_samples, device=verts.device) * cdf[:, -1:]
    face_idx = torch.searchsorted(cdf, r, right=True).clamp_max(areas.shape[1] - 1)                # B, S
    u = torch.rand(B, num_samples, 2, device=verts.device)
    su = u[..., 0].sqrt()
    return face_idx, torch.stack((1 - su, su * (1 - u[..., 1]), su * u[..., 1]), dim=-1)


def points_from_samples(verts, faces, face_idx, w):
    """pts = w0 v0 + w1 v1 + w2 v2 of the drawn faces, (B,S,3); differentiable in verts."""
    idx = faces.long()
    corner = lambda k: torch.gather(verts, 1, torch.gather(idx[:, :, k], 1, face_idx)[:, :, None].expand(-1, -1, 3))
    return w[..., 0:1] * corner(0) + w[..., 1:2] * corner(1) + w[..., 2:3] * corner(2)


def sample_points_from_meshes(verts, faces, num_samples):
    """Area-weighted uniform surface samples, (B, num_samples, 3) (pytorch3d.ops.sample_points_from_meshes semantics)."""
    face_idx, w = sample_faces_and_weights(verts, faces, num_samples)
    return points_from_samples(verts, faces, face_idx, w)


def chamfer_single_way(x, y, chunk=16):
    """mean over x-points of the squared distance to their nearest y-point, averaged over the batch
    (model/util/chamfer.py chamfer_distance_single_way with the default mean reductions)."""
    total = x.new_zeros(())
    for s in range(0, x.shape[0], chunk):
        # exact differences like pytorch3d's knn_points (the default cdist goes through |x|^2 + |y|^2 - 2 x.y, whose
        # cancellation error (~1e-7) reorders near-tied neighbours)
        d = torch.cdist(x[s:s + chunk], y[s:s + chunk], compute_mode='donot_use_mm_for_euclid_dist').pow(2).min(dim=2)[0]     # b, Nx
        total = total + d.mean(1).sum()
    return total / x.shape[0]


class CanonicalMesh(nn.Module):

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.mean_v, self.faces, self.symm_rots = self.init_shape()
        self.num_verts, self.num_faces = self.mean_v.shape[0], self.faces.shape[0]
        if getattr(opts, 'surface_texture', False):
            raise NotImplementedError('surface textures are off in every shipped config and outside the hot path')
        self.texture_type = 'vertex'

    def get_texture(self, pred_v, faces, imatch, img):
        return F.grid_sample(img, imatch.permute(0, 2, 1)[:, None], align_corners=False)[:, :, 0].permute(0, 2, 1)

    def compute_symmetry_loss(self, pred_v, faces, npts=10000):
        """mesh.py:53-62 of the reference.  On the GPU the sample reconstruction, the rotation and the 1-NN search of the
        one-way chamfer distance run as ONE native kernel (ops/symmetry_nn.py); the random draws are the same torch calls
        in the same order as `compute_symmetry_loss_reference`, so both produce the same value for the same seed."""
        if not pred_v.is_cuda:
            return self.compute_symmetry_loss_reference(pred_v, faces, npts)
        from ...ops.symmetry_nn import symmetry_nn
        k, bsz = self.symm_rots.shape[0], pred_v.shape[0]
        v = pred_v[:, None].repeat(1, k, 1, 1).reshape(k * bsz, self.num_verts, 3)
        f = faces[:, None].repeat(1, k, 1, 1).reshape(k * bsz, self.num_faces, 3)
        with torch.no_grad():
            face_idx, w = (getattr(self, 'sampler', None) or sample_faces_and_weights)(v, f, npts)
        if getattr(self, '_faces_i32', None) is None or self._faces_i32.device != pred_v.device:
            self._faces_i32 = self.faces.detach().to(pred_v.device, torch.int32).contiguous()
        dist, _ = symmetry_nn(pred_v, self._faces_i32, face_idx, w, self.symm_rots)
        return dist.mean(1).mean(0)

    def compute_symmetry_loss_reference(self, pred_v, faces, npts=10000):
        """The reference's statements op by op (parity reference of the fused kernel; CPU-capable)."""
        k, bsz = self.symm_rots.shape[0], pred_v.shape[0]
        v = pred_v[:, None].repeat(1, k, 1, 1).reshape(k * bsz, self.num_verts, 3)
        f = faces[:, None].repeat(1, k, 1, 1).reshape(k * bsz, self.num_faces, 3)
        face_idx, w = (getattr(self, 'sampler', None) or sample_faces_and_weights)(v, f, npts)
        pts = points_from_samples(v, f, face_idx, w)
        rots = self.symm_rots[None].repeat(bsz, 1, 1, 1).reshape(k * bsz, 3, 3)
        # pts.bmm(rots) of the reference (mesh.py:61), written element-wise: cuBLAS runs a (10000 x 3) x (3 x 3) batch as
        # one gemv launch PER MESH (128 launches, 14 ms per step at B = 64 on a B200)
        rotated = (pts[:, :, :, None] * rots[:, None, :, :]).sum(2)
        return chamfer_single_way(v, rotated)

    def _load_prior(self, path):
        if path == 'synthetic:uv1280':     # the 1280-vertex / 2556-face sphere of BASELINE.json (benchmarks, tests)
            return synthetic.uv_sphere()
        if os.path.exists(path):
            return read_obj(path)
        if not _flags.synthetic_weights_allowed():
            raise FileNotFoundError('shape prior %s not found (set SCP_SYNTHETIC_WEIGHTS=1 to fall back to the packaged '
                                    'copy of the category prior, tests / benchmarks only)' % path)
        cat = os.path.splitext(os.path.basename(path))[0]
        return synthetic.load_prior(cat, normalise=False)     # packaged copy of config/<cat>_wild6d/<cat>.obj

    def init_shape(self):
        opts = self.opts
        if opts.shape_prior:
            v, f = self._load_prior(opts.shape_prior_path)
            verts, faces = torch.from_numpy(v).float(), torch.from_numpy(f)
            verts = verts - verts.mean(0)
            verts = verts / verts.abs().max()
            verts = verts * torch.tensor([float(s) for s in opts.init_scale])
            learn = opts.prior_deform
        else:
            v, f = synthetic.icosphere(opts.subdivide)
            verts, faces = torch.from_numpy(v).float(), torch.from_numpy(f)
            verts = verts * torch.tensor([getattr(opts, 'x_scale', 1.), getattr(opts, 'y_scale', 1.), getattr(opts, 'z_scale', 1.)])
            learn = True
        if opts.symmetry_idx == 0:
            symm = get_symm_rots(17)
        elif opts.symmetry_idx == 1:
            symm = torch.stack([torch.eye(3), torch.diag(torch.tensor([-1., 1., 1.]))])
        else:
            symm = torch.eye(3)[None]
        return (nn.Parameter(verts.float(), requires_grad=learn), nn.Parameter(faces.long(), requires_grad=False),
                nn.Parameter(symm, requires_grad=False))
