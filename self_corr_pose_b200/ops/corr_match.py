"""Fused Correspondence.match operator (forward + backward on the sm_100a kernels).

Replaces the torch op chain of model/module/correspondence.py:42-53 of the reference
(bmm, mask, softmax(dim=1), softmax(dim=2), meshgrid.bmm, broadcast-multiply-sum).
"""
import os

import torch
from torch.autograd import Function

from .. import _lib


def check_shapes(img_feat, mesh_feat, mask_down, pred_v, meshgrid, hf, wf):
    """(B, C, P, N) after checking every buffer against the element count the kernels index."""
    B, C, P = img_feat.shape
    N = mesh_feat.shape[1]
    if P != hf * wf:
        raise ValueError('corr_match: %d pixel features for a %d x %d feature map' % (P, hf, wf))
    _lib.expect_numel('corr_match', mesh_feat=(mesh_feat, B * N * C), mask_down=(mask_down, B * P),
                      pred_v=(pred_v, B * N * 3), meshgrid=(meshgrid, 2 * P))
    return B, C, P, N


class CorrMatchFunction(Function):
    """(img_feat[B,C,P], mesh_feat[B,N,C], mask_down[B,P], pred_v[B,N,3], meshgrid[2,P]) ->
    (pointcorr_full[B,P,N] | None, pointcorr_pool[B,P/4,N] | None, match[B,P,3], imatch[B,2,N],
     A_pool[B,2,N] | None)   -- A_pool = pooled_meshgrid . softmax(tau * pointcorr_pool, dim=pixels), produced with
    pointcorr_pool (the "grid.bmm(pointcorr_mesh)" factor of the pre-training cycle loss)"""

    @staticmethod
    def forward(ctx, img_feat, mesh_feat, mask_down, pred_v, meshgrid, tau, hf, wf, want_full, want_pool):
        B, C, P, N = check_shapes(img_feat, mesh_feat, mask_down, pred_v, meshgrid, hf, wf)
        dev = img_feat.device
        img_feat = img_feat.detach().float().contiguous()
        mesh_feat = mesh_feat.detach().float().contiguous()
        mask_down = mask_down.detach().float().contiguous()
        pred_v = pred_v.detach().float().contiguous()
        meshgrid = meshgrid.detach().float().contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        pc_full = torch.empty(B, P, N, **f32) if want_full else None
        pc_pool = torch.empty(B, P // 4, N, **f32) if want_pool else None
        match = torch.empty(B, P, 3, **f32)
        imatch = torch.empty(B, 2, N, **f32)
        rsum = torch.empty(B, P, **f32)
        csum = torch.empty(B, N, **f32)
        A_pool = torch.empty(B, 2, N, **f32) if want_pool else None
        csum_pool = torch.empty(B, N, **f32) if want_pool else None
        L = _lib.lib()
        ws_bytes = L.scp_corr_workspace_bytes(B, hf, wf, N)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = L.scp_corr_match_forward(
                _lib.ptr(img_feat), _lib.ptr(mesh_feat), _lib.ptr(mask_down), _lib.ptr(pred_v), _lib.ptr(meshgrid),
                float(tau), B, hf, wf, N, C, _lib.ptr(pc_full), _lib.ptr(pc_pool), _lib.ptr(match),
                _lib.ptr(imatch), _lib.ptr(rsum), _lib.ptr(csum), _lib.ptr(A_pool), _lib.ptr(csum_pool), _lib.ptr(ws),
                ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_corr_match_forward')
        ctx.save_for_backward(img_feat, mesh_feat, mask_down, pred_v, meshgrid, match, imatch, rsum, csum, A_pool, csum_pool)
        ctx.geom = (float(tau), B, hf, wf, N, C)
        ctx.set_materialize_grads(False)
        return pc_full, pc_pool, match, imatch, A_pool

    @staticmethod
    def backward(ctx, g_full, g_pool, g_match, g_imatch, g_A):
        img_feat, mesh_feat, mask_down, pred_v, meshgrid, match, imatch, rsum, csum, A_pool, csum_pool = ctx.saved_tensors
        tau, B, hf, wf, N, C = ctx.geom
        dev = img_feat.device

        def prep(g, like_shape):
            if g is None:
                return None
            return g.float().contiguous()
        g_full, g_pool, g_A = prep(g_full, None), prep(g_pool, None), prep(g_A, None)
        g_match = prep(g_match, None) if g_match is not None else torch.zeros_like(match)
        g_imatch = prep(g_imatch, None) if g_imatch is not None else torch.zeros_like(imatch)
        g_img = torch.empty_like(img_feat)
        g_mesh = torch.empty_like(mesh_feat)
        ws_bytes = _lib.lib().scp_corr_backward_workspace_bytes(B, hf, wf, N)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_corr_match_backward(
                _lib.ptr(img_feat), _lib.ptr(mesh_feat), _lib.ptr(mask_down), _lib.ptr(pred_v), _lib.ptr(meshgrid),
                tau, B, hf, wf, N, C, _lib.ptr(match), _lib.ptr(imatch), _lib.ptr(rsum), _lib.ptr(csum),
                _lib.ptr(g_match), _lib.ptr(g_imatch), _lib.ptr(g_pool), _lib.ptr(g_full), _lib.ptr(A_pool),
                _lib.ptr(csum_pool), _lib.ptr(g_A), _lib.ptr(g_img), _lib.ptr(g_mesh), _lib.ptr(ws), ws_bytes,
                _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_corr_match_backward')
        return g_img, g_mesh, None, None, None, None, None, None, None, None


def corr_match(img_feat, mesh_feat, mask_down, pred_v, meshgrid, tau, hf, wf, want_full=False, want_pool=True):
    """The kernels' soft-maxes use the fixed reference point tau * 1: `img_feat` (over C) and `mesh_feat` (over C) must be
    L2-normalised, as the reference's encoder produces them (encoder.py:36,45), and 0 < tau <= 40.  SCP_DEBUG_CHECKS=1
    verifies the norms on every call (one host synchronisation)."""
    if not img_feat.is_cuda:
        raise TypeError('corr_match supports only CUDA tensors (no CPU path)')
    if not 0. < float(tau) <= 40.:
        raise ValueError('corr_match: tau = %g outside (0, 40]' % float(tau))
    if os.environ.get('SCP_DEBUG_CHECKS') == '1':
        worst = max(float(img_feat.detach().norm(dim=1).max()), float(mesh_feat.detach().norm(dim=2).max()))
        if worst > 1. + 1e-3:
            raise ValueError('corr_match: features are not L2-normalised (largest norm %.4f): the fixed-reference-point '
                             'soft-max would overflow / lose accuracy' % worst)
    return CorrMatchFunction.apply(img_feat, mesh_feat, mask_down, pred_v, meshgrid, tau, hf, wf, want_full,
                                   want_pool)
