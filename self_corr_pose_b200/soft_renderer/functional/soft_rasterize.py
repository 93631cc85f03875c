"""SoftRasterizeFunction -- the reference's autograd operator, routed to the sm_100a C ABI.

Signature and semantics follow third-party/softras/soft_renderer/functional/soft_rasterize.py:9-117
(SoftRasterizeFunction.forward/backward, soft_rasterize); the two pybind calls
forward_soft_rasterize / backward_soft_rasterize (cuda/soft_rasterize_cuda.cpp:59-132) become
scp_softras_forward / scp_softras_backward (include/scp_b200.h).  Unlike the reference the
kernels run on torch's current stream.
"""
import math

import torch
from torch.autograd import Function

from ... import _lib

FUNC_DIST = {'hard': 0, 'barycentric': 1, 'euclidean': 2}
FUNC_RGB = {'hard': 0, 'softmax': 1}
FUNC_ALPHA = {'hard': 0, 'sum': 1, 'prod': 2}
FUNC_TEX = {'surface': 0, 'vertex': 1}


def _scalars(ctx):
    return (ctx.batch_size, ctx.num_faces, ctx.texture_size, ctx.image_size, ctx.near, ctx.far, ctx.eps,
            ctx.sigma_val, ctx.func_dist_type, ctx.dist_eps, ctx.gamma_val, ctx.func_rgb_type,
            ctx.func_alpha_type, ctx.texture_type, int(bool(ctx.fill_back)))


def _workspace(B, nf, device):
    n = _lib.lib().scp_softras_workspace_bytes(B, nf)
    return torch.empty(n, dtype=torch.uint8, device=device), n


class SoftRasterizeFunction(Function):

    @staticmethod
    def forward(ctx, face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4,
                aggr_func_rgb='softmax', aggr_func_alpha='prod', texture_type='surface'):
        if not face_vertices.is_cuda:
            raise TypeError('Rasterize module supports only cuda Tensors')
        ctx.image_size = image_size
        ctx.background_color = background_color
        ctx.near, ctx.far, ctx.eps = float(near), float(far), float(eps)
        ctx.sigma_val, ctx.gamma_val = float(sigma_val), float(gamma_val)
        ctx.func_dist_type = FUNC_DIST[dist_func]
        ctx.dist_eps = math.log(1. / dist_eps - 1.)
        ctx.func_rgb_type = FUNC_RGB[aggr_func_rgb]
        ctx.func_alpha_type = FUNC_ALPHA[aggr_func_alpha]
        ctx.texture_type = FUNC_TEX[texture_type]
        ctx.fill_back = fill_back

        ctx.batch_size, ctx.num_faces = face_vertices.shape[:2]
        B, nf = ctx.batch_size, ctx.num_faces
        dev = face_vertices.device
        ctx.in_shapes = (face_vertices.shape, textures.shape)
        face_vertices = face_vertices.detach().float().reshape(B, nf, 3, 3).contiguous().clone()
        textures = textures.detach().float().reshape(B, nf, -1, 3).contiguous().clone()
        ctx.texture_size = textures.shape[2]

        faces_info = torch.zeros(B, nf, 27, dtype=torch.float32, device=dev)
        aggrs_info = torch.zeros(B, 2, image_size, image_size, dtype=torch.float32, device=dev)
        soft_colors = torch.ones(B, 4, image_size, image_size, dtype=torch.float32, device=dev)
        for k in range(3):
            if background_color[k] != 1:
                soft_colors[:, k] *= background_color[k]

        ws, ws_bytes = _workspace(B, nf, dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_softras_forward(
                _lib.ptr(face_vertices), _lib.ptr(textures), _lib.ptr(faces_info), _lib.ptr(aggrs_info),
                _lib.ptr(soft_colors), *_scalars(ctx), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_softras_forward')
        ctx.save_for_backward(face_vertices, textures, soft_colors, faces_info, aggrs_info)
        return soft_colors

    @staticmethod
    def backward(ctx, grad_soft_colors):
        face_vertices, textures, soft_colors, faces_info, aggrs_info = ctx.saved_tensors
        dev = face_vertices.device
        grad_faces = torch.zeros_like(face_vertices)
        grad_textures = torch.zeros_like(textures)
        grad_soft_colors = grad_soft_colors.float().contiguous()
        ws, ws_bytes = _workspace(ctx.batch_size, ctx.num_faces, dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_softras_backward(
                _lib.ptr(face_vertices), _lib.ptr(textures), _lib.ptr(soft_colors), _lib.ptr(faces_info),
                _lib.ptr(aggrs_info), _lib.ptr(grad_faces), _lib.ptr(grad_textures), _lib.ptr(grad_soft_colors),
                *_scalars(ctx), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_softras_backward')
        return (grad_faces.reshape(ctx.in_shapes[0]), grad_textures.reshape(ctx.in_shapes[1])) + (None,) * 13


class SoftRasterizeDualFunction(Function):
    """Depth render (softmax RGB of `textures_soft`) and NOCS map (hard RGB of `textures_hard`) in one traversal
    (scp_softras_forward_dual).  Returns (soft_colors_soft, soft_colors_hard); only the first is differentiable
    (w.r.t. face_vertices and textures_soft, through scp_softras_backward) -- the hard render's useful gradient is
    exactly zero on this path."""

    @staticmethod
    def forward(ctx, face_vertices, textures_soft, textures_hard, image_size, background_soft, background_hard, near,
                far, fill_back, eps, sigma_val, dist_eps, gamma_val):
        if not face_vertices.is_cuda:
            raise TypeError('Rasterize module supports only cuda Tensors')
        ctx.image_size = image_size
        ctx.near, ctx.far, ctx.eps = float(near), float(far), float(eps)
        ctx.sigma_val, ctx.gamma_val = float(sigma_val), float(gamma_val)
        ctx.func_dist_type, ctx.func_rgb_type, ctx.func_alpha_type = FUNC_DIST['euclidean'], FUNC_RGB['softmax'], \
            FUNC_ALPHA['prod']
        ctx.dist_eps = math.log(1. / dist_eps - 1.)
        ctx.texture_type, ctx.texture_size, ctx.fill_back = FUNC_TEX['vertex'], 3, fill_back
        ctx.batch_size, ctx.num_faces = face_vertices.shape[:2]
        B, nf = ctx.batch_size, ctx.num_faces
        dev = face_vertices.device
        ctx.in_shapes = (face_vertices.shape, textures_soft.shape)
        face_vertices = face_vertices.detach().float().reshape(B, nf, 3, 3).contiguous()
        textures_soft = textures_soft.detach().float().reshape(B, nf, 3, 3).contiguous()
        textures_hard = textures_hard.detach().float().reshape(B, nf, 3, 3).contiguous()
        faces_info = torch.zeros(B, nf, 27, dtype=torch.float32, device=dev)
        outs = []
        for bg in (background_soft, background_hard):
            aggrs = torch.zeros(B, 2, image_size, image_size, dtype=torch.float32, device=dev)
            cols = torch.ones(B, 4, image_size, image_size, dtype=torch.float32, device=dev)
            for k in range(3):
                if bg[k] != 1:
                    cols[:, k] *= bg[k]
            outs.append((aggrs, cols))
        ws, ws_bytes = _workspace(B, nf, dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_softras_forward_dual(
                _lib.ptr(face_vertices), _lib.ptr(textures_soft), _lib.ptr(textures_hard), _lib.ptr(faces_info),
                _lib.ptr(outs[0][0]), _lib.ptr(outs[0][1]), _lib.ptr(outs[1][0]), _lib.ptr(outs[1][1]), B, nf,
                image_size, ctx.near, ctx.far, ctx.eps, ctx.sigma_val, ctx.dist_eps, ctx.gamma_val,
                int(bool(fill_back)), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_softras_forward_dual')
        ctx.save_for_backward(face_vertices, textures_soft, outs[0][1], faces_info, outs[0][0])
        ctx.mark_non_differentiable(outs[1][1])
        return outs[0][1], outs[1][1]

    @staticmethod
    def backward(ctx, grad_soft_colors, _grad_hard):
        grads = SoftRasterizeFunction.backward(ctx, grad_soft_colors)
        return (grads[0], grads[1]) + (None,) * 11


def soft_rasterize_dual(face_vertices, textures_soft, textures_hard, image_size=256, background_soft=(1, 1, 1),
                        background_hard=(0, 0, 0), near=1, far=100, fill_back=True, eps=1e-3, sigma_val=1e-4,
                        dist_eps=1e-4, gamma_val=1e-4):
    return SoftRasterizeDualFunction.apply(face_vertices, textures_soft, textures_hard, image_size,
                                           list(background_soft), list(background_hard), near, far, fill_back, eps,
                                           sigma_val, dist_eps, gamma_val)


def soft_rasterize(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                   fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4,
                   gamma_val=1e-4, aggr_func_rgb='softmax', aggr_func_alpha='prod', texture_type='surface'):
    if not face_vertices.is_cuda:
        raise TypeError('Rasterize module supports only cuda Tensors')
    return SoftRasterizeFunction.apply(face_vertices, textures, image_size, background_color, near, far,
                                       fill_back, eps, sigma_val, dist_func, dist_eps, gamma_val,
                                       aggr_func_rgb, aggr_func_alpha, texture_type)
