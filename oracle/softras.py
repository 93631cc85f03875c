"""ctypes front-end of the SoftRas CPU checkers -- TEST INFRASTRUCTURE ONLY.

  * `forward` / `backward`      -> oracle/liboracle_softras.so  (our C restatement,
                                   oracle/softras_oracle.c)
  * `ref_forward`/`ref_backward`-> oracle/_ref/libsoftras_ref_cpu.so (the reference's own
                                   kernel source compiled for the host, oracle/Makefile)

Argument meaning follows forward_soft_rasterize / backward_soft_rasterize of
third-party/softras/soft_renderer/cuda/soft_rasterize_cuda.cpp:59-132 and the buffer
initialisation of functional/soft_rasterize.py:35,47-53,88-89 (soft_colors pre-filled with the
background colour; dist_eps already transformed to ln(1/dist_eps - 1)).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DIST = {'hard': 0, 'barycentric': 1, 'euclidean': 2}
RGB = {'hard': 0, 'softmax': 1}
ALPHA = {'hard': 0, 'sum': 1, 'prod': 2}
TEX = {'surface': 0, 'vertex': 1}

_fp = ctypes.POINTER(ctypes.c_float)
_SCALARS = [ctypes.c_int] * 4 + [ctypes.c_float] * 4 + [ctypes.c_int, ctypes.c_float, ctypes.c_float] + \
    [ctypes.c_int] * 4


def build():
    """(Re)build the checker libraries with oracle/Makefile."""
    subprocess.run(['make', '-s', '-C', _HERE], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _load(path, prefix, extra_int):
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    fwd = getattr(lib, prefix + 'forward')
    bwd = getattr(lib, prefix + 'backward')
    fwd.argtypes = [_fp] * 5 + _SCALARS + [ctypes.c_int] * extra_int
    bwd.argtypes = [_fp] * 8 + _SCALARS + [ctypes.c_int] * extra_int
    fwd.restype = bwd.restype = ctypes.c_int
    return fwd, bwd


_cache = {}


def _get(kind):
    if kind not in _cache:
        if kind == 'oracle':
            _cache[kind] = _load(os.path.join(_HERE, 'liboracle_softras.so'), 'scp_oracle_softras_', 1)
        elif kind == 'oracle_fma':
            _cache[kind] = _load(os.path.join(_HERE, 'liboracle_softras_fma.so'), 'scp_oracle_softras_', 1)
        else:
            _cache[kind] = _load(os.path.join(_HERE, '_ref', 'libsoftras_ref_cpu.so'), 'scp_ref_softras_', 0)
    return _cache[kind]


def have_ref():
    return os.path.exists(os.path.join(_HERE, '_ref', 'libsoftras_ref_cpu.so')) or \
        os.path.exists('/root/reference/third-party/softras')


def _p(a):
    return a.ctypes.data_as(_fp)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _scalars(B, nf, T, image_size, near, far, eps, sigma_val, dist_func, dist_eps, gamma_val, aggr_func_rgb,
             aggr_func_alpha, texture_type, fill_back):
    return (B, nf, T, image_size, near, far, eps, sigma_val, DIST[dist_func],
            float(np.log(1. / dist_eps - 1.)), gamma_val, RGB[aggr_func_rgb], ALPHA[aggr_func_alpha],
            TEX[texture_type], int(bool(fill_back)))


def _forward(kind, face_vertices, textures, image_size=256, background_color=(0, 0, 0), near=1, far=100,
             fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4,
             aggr_func_rgb='softmax', aggr_func_alpha='prod', texture_type='surface', nthreads=0):
    faces = _f32(face_vertices).reshape(face_vertices.shape[0], face_vertices.shape[1], 9)
    B, nf = faces.shape[:2]
    tex = _f32(textures).reshape(B, nf, -1, 3)
    T = tex.shape[2]
    # one padded face of texture: the reference can read a texel past the end (see softras_oracle.c)
    tex_pad = np.concatenate([tex.reshape(-1), np.ones(3 * T, np.float32)])
    faces_info = np.zeros((B, nf, 27), np.float32)
    aggrs_info = np.zeros((B, 2, image_size, image_size), np.float32)
    soft_colors = np.ones((B, 4, image_size, image_size), np.float32)
    for k in range(3):
        soft_colors[:, k] *= background_color[k]
    fwd, _ = _get(kind)
    args = [_p(faces), _p(tex_pad), _p(faces_info), _p(aggrs_info), _p(soft_colors)] + \
        list(_scalars(B, nf, T, image_size, near, far, eps, sigma_val, dist_func, dist_eps, gamma_val,
                      aggr_func_rgb, aggr_func_alpha, texture_type, fill_back))
    if kind.startswith('oracle'):
        args.append(nthreads)
    rc = fwd(*args)
    assert rc == 0
    return soft_colors, faces_info, aggrs_info


def _backward(kind, face_vertices, textures, soft_colors, faces_info, aggrs_info, grad_soft_colors,
              image_size=256, background_color=(0, 0, 0), near=1, far=100, fill_back=True, eps=1e-3,
              sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb='softmax',
              aggr_func_alpha='prod', texture_type='surface', nthreads=0):
    faces = _f32(face_vertices).reshape(face_vertices.shape[0], face_vertices.shape[1], 9)
    B, nf = faces.shape[:2]
    tex = _f32(textures).reshape(B, nf, -1, 3)
    T = tex.shape[2]
    tex_pad = np.concatenate([tex.reshape(-1), np.ones(3 * T, np.float32)])
    grad_faces = np.zeros((B, nf, 9), np.float32)
    grad_textures = np.zeros((B, nf, T, 3), np.float32)
    g = _f32(grad_soft_colors)
    _, bwd = _get(kind)
    args = [_p(faces), _p(tex_pad), _p(_f32(soft_colors)), _p(_f32(faces_info)), _p(_f32(aggrs_info)),
            _p(grad_faces), _p(grad_textures), _p(g)] + \
        list(_scalars(B, nf, T, image_size, near, far, eps, sigma_val, dist_func, dist_eps, gamma_val,
                      aggr_func_rgb, aggr_func_alpha, texture_type, fill_back))
    if kind.startswith('oracle'):
        args.append(nthreads)
    rc = bwd(*args)
    assert rc == 0
    return grad_faces.reshape(B, nf, 3, 3), grad_textures


def forward(*a, fma=False, **k):
    """Our restatement. Returns (soft_colors[B,4,is,is], faces_info[B,nf,27], aggrs_info[B,2,is,is]).
    fma=True runs the build with FMA contraction (same source, nvcc-like rounding)."""
    return _forward('oracle_fma' if fma else 'oracle', *a, **k)


def backward(*a, fma=False, **k):
    """Our restatement. Returns (grad_faces[B,nf,3,3], grad_textures[B,nf,T,3])."""
    return _backward('oracle_fma' if fma else 'oracle', *a, **k)


def ref_forward(*a, **k):
    k.pop('nthreads', None)
    return _forward('ref', *a, **k)


def ref_backward(*a, **k):
    k.pop('nthreads', None)
    return _backward('ref', *a, **k)
