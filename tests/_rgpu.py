"""Loader for "R-GPU": the reference's own legacy SoftRas CUDA operator recompiled for sm_100a
(baseline/build_ref_gpu.py -> baseline/_ref/soft_rasterize.so; git-ignored, travels to the GPU box).  Test /
measurement infrastructure only.  The two pybind entry points have the reference's signatures
(third-party/softras/soft_renderer/cuda/soft_rasterize_cuda.cpp:59-132)."""
import importlib.util
import math
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'baseline', '_ref', 'soft_rasterize.so')
_mod = None

DIST = {'hard': 0, 'barycentric': 1, 'euclidean': 2}
RGB = {'hard': 0, 'softmax': 1}
ALPHA = {'hard': 0, 'sum': 1, 'prod': 2}
TEX = {'surface': 0, 'vertex': 1}


def available():
    return os.path.exists(SO)


def module():
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location('soft_rasterize', SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def scalars(image_size, sigma_val, gamma_val, aggr_func_rgb, texture_type, dist_func='euclidean', dist_eps=1e-4,
            aggr_func_alpha='prod', near=1., far=100., eps=1e-3, fill_back=True, **_):
    # functional/soft_rasterize.py:22-35 of the reference's SoftRas package
    return (image_size, near, far, eps, sigma_val, DIST[dist_func], math.log(1. / dist_eps - 1.), gamma_val,
            RGB[aggr_func_rgb], ALPHA[aggr_func_alpha], TEX[texture_type], fill_back)


def forward(fv, tex, background_color=(0, 0, 0), **kw):
    """Buffer set-up of SoftRasterizeFunction.forward (functional/soft_rasterize.py:37-58) + the legacy launch.
    The legacy kernels run on the legacy default stream: synchronise around the call."""
    m = module()
    B, nf = fv.shape[:2]
    is_ = kw['image_size']
    dev = fv.device
    faces_info = torch.zeros(B, nf, 9 * 3, device=dev)
    aggrs_info = torch.zeros(B, 2, is_, is_, device=dev)
    soft_colors = torch.ones(B, 4, is_, is_, device=dev)
    for c in range(3):
        soft_colors[:, c] *= background_color[c]
    torch.cuda.synchronize()
    faces_info, aggrs_info, soft_colors = m.forward_soft_rasterize(fv.contiguous(), tex.contiguous(), faces_info,
                                                                   aggrs_info, soft_colors, *scalars(**kw))
    torch.cuda.synchronize()
    return soft_colors, faces_info, aggrs_info


def backward(fv, tex, soft_colors, faces_info, aggrs_info, grad_soft_colors, **kw):
    m = module()
    grad_faces = torch.zeros_like(fv)
    grad_textures = torch.zeros_like(tex)
    torch.cuda.synchronize()
    grad_faces, grad_textures = m.backward_soft_rasterize(fv.contiguous(), tex.contiguous(), soft_colors, faces_info,
                                                          aggrs_info, grad_faces, grad_textures,
                                                          grad_soft_colors.contiguous(), *scalars(**kw))
    torch.cuda.synchronize()
    return grad_faces, grad_textures
