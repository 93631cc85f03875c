"""Pure-torch geometry helpers of the SoftRas front-end (device agnostic).

Mirrors third-party/softras/soft_renderer/functional/{face_vertices.py:4-22,
vertex_normals.py:4-37, look_at.py:6-62, orthogonal.py:4-16, perspective.py,
ambient_lighting.py:7-18, directional_lighting.py:7-31}.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def face_vertices(vertices, faces):
    """[B,V,3] x [B,F,3] (int) -> [B,F,3,3] gather of per-face vertex attributes."""
    if vertices.ndimension() != 3 or faces.ndimension() != 3:
        raise AssertionError('vertices and faces must be 3-dimensional')
    if vertices.shape[0] != faces.shape[0] or vertices.shape[2] != 3 or faces.shape[2] != 3:
        raise AssertionError('shape mismatch between vertices and faces')
    B, nf = faces.shape[:2]
    idx = faces.long().reshape(B, nf * 3, 1).expand(-1, -1, 3)
    return torch.gather(vertices, 1, idx).reshape(B, nf, 3, 3)


def vertex_normals(vertices, faces):
    """Area-weighted vertex normals [B,V,3] (sum of incident face cross products, normalised)."""
    B, nv = vertices.shape[:2]
    fv = face_vertices(vertices, faces)
    v0, v1, v2 = fv[:, :, 0], fv[:, :, 1], fv[:, :, 2]
    normals = torch.zeros(B * nv, 3, dtype=vertices.dtype, device=vertices.device)
    flat = (faces.long() + (torch.arange(B, device=faces.device) * nv)[:, None, None]).reshape(-1, 3)
    corners = ((1, v2 - v1, v0 - v1), (2, v0 - v2, v1 - v2), (0, v1 - v0, v2 - v0))
    for k, a, b in corners:
        normals.index_add_(0, flat[:, k], torch.cross(a, b, dim=-1).reshape(-1, 3))
    return F.normalize(normals, eps=1e-6, dim=1).reshape(B, nv, 3)


_CONST_CACHE = {}


def _const(values, device):
    """Small constant vectors given as python lists (eye, up, light colours): uploaded once per device and cached,
    so that no host->device copy happens inside a CUDA-graph capture."""
    key = (tuple(float(v) for v in values), str(device))
    if key not in _CONST_CACHE:
        _CONST_CACHE[key] = torch.tensor([float(v) for v in values], dtype=torch.float32, device=device)
    return _CONST_CACHE[key]


def _as_vec(x, device, B):
    if isinstance(x, (list, tuple)):
        x = _const(x, device)
    elif isinstance(x, np.ndarray):
        x = torch.from_numpy(x).to(device)
    else:
        x = x.to(device)
    return x[None, :].repeat(B, 1) if x.ndimension() == 1 else x


def look_at(vertices, eye, at=(0, 0, 0), up=(0, 1, 0)):
    """World -> camera frame for a camera at `eye` looking at `at` (look_at.py:6-62)."""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    B, dev = vertices.shape[0], vertices.device
    eye, at, up = _as_vec(eye, dev, B), _as_vec(at, dev, B), _as_vec(up, dev, B)
    z = F.normalize(at - eye, eps=1e-5)
    x = F.normalize(torch.cross(up, z, dim=-1), eps=1e-5)
    y = F.normalize(torch.cross(z, x, dim=-1), eps=1e-5)
    rot = torch.stack((x, y, z), dim=1)
    if vertices.shape != eye.shape:
        eye = eye[:, None, :]
    return torch.matmul(vertices - eye, rot.transpose(1, 2))


def orthogonal(vertices, scale):
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    return torch.stack((vertices[:, :, 0] * scale, vertices[:, :, 1] * scale, vertices[:, :, 2]), dim=2)


def perspective(vertices, angle=30.):
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    width = math.tan(math.radians(angle))
    z = vertices[:, :, 2]
    return torch.stack((vertices[:, :, 0] / z / width, vertices[:, :, 1] / z / width, z), dim=2)


def _as_row(x, device):
    if isinstance(x, (list, tuple)):
        x = _const(x, device)
    elif isinstance(x, np.ndarray):
        x = torch.from_numpy(x).float().to(device)
    return x[None, :] if x.ndimension() == 1 else x


def ambient_lighting(light, light_intensity=0.5, light_color=(1, 1, 1)):
    light += light_intensity * _as_row(light_color, light.device)[:, None, :]
    return light


def directional_lighting(light, normals, light_intensity=0.5, light_color=(1, 1, 1), light_direction=(0, 1, 0)):
    color = _as_row(light_color, light.device)
    direction = _as_row(light_direction, light.device)
    cosine = F.relu(torch.sum(normals * direction, dim=2))
    light += light_intensity * (color[:, None, :] * cosine[:, :, None])
    return light
