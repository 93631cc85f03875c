"""Fused image-space loss terms (forward + backward on the sm_100a kernels of csrc/scp_loss.cu).

Replaces the torch op chains of the reference's model/util/loss_utils.py:236-244 (compute_mask_loss), :246-252
(compute_texture_loss), :273-284 (compute_depth_loss) and :317-320 (compute_match_loss, including the nearest
upsampling of `match`, model/module/correspondence.py:71).  The render inputs are the raw (B,4,H,W) SoftRas outputs:
their channels are read in place and the gradients are written straight into (B,4,H,W) tensors, so no slicing /
zero-filling glue runs between the losses and the SoftRas backward.
"""
import ctypes

import torch
from torch.autograd import Function

from .. import _lib

_NIN, _NOUT = 11, 5


def check_shapes(r_depth, r_tex, match_lr, img, mask, depth, r_nocs, hf, wf):
    """(B, H, W) after checking every map and the match buffer against what the kernels index."""
    B, ch, H, W = r_depth.shape
    if ch != 4 or r_tex.shape != r_depth.shape or r_nocs.shape != r_depth.shape:
        raise ValueError('image losses: renders must be three (B,4,H,W) maps, got %s %s %s'
                         % (tuple(r_depth.shape), tuple(r_tex.shape), tuple(r_nocs.shape)))
    _lib.expect_numel('image_losses', match_lr=(match_lr, B * hf * wf * 3), img=(img, B * 3 * H * W),
                      mask=(mask, B * H * W), depth=(depth, B * H * W))
    for name, t in (('img', img), ('mask', mask), ('depth', depth)):
        if t is not None and (t.shape[0] != B or tuple(t.shape[-2:]) != (H, W)):
            raise ValueError('image losses: `%s` of shape %s for (%d, ..., %d, %d) renders' % (name, tuple(t.shape), B, H, W))
    return B, H, W


def _view(t, channels, H, W):
    """(device pointer, batch stride) of a (B,[C,]H,W) view with contiguous planes."""
    if t is None:
        return None, 0
    if not t.is_cuda or t.dtype != torch.float32:
        raise _lib.ScpNativeError('image losses: expected CUDA fp32 maps (no CPU path)')
    ok = t.stride(-1) == 1 and t.stride(-2) == W and (channels == 1 or t.stride(1) == H * W)
    if not ok:
        raise _lib.ScpNativeError('image losses: map planes must be contiguous (got strides %s)' % (t.stride(),))
    return t.data_ptr(), t.stride(0)


def _pack(views):
    ptrs = (ctypes.c_void_p * len(views))(*[v[0] for v in views])
    strides = (ctypes.c_longlong * len(views))(*[v[1] for v in views])
    return ptrs, strides


class ImageLossFunction(Function):
    """(r_depth[B,4,H,W], r_tex[B,4,H,W], match_lr[B,P,3]; img, mask, depth, r_nocs) -> losses[B,4].

    r_depth: depth render (channel 2 = depth, channel 3 = alpha = mask_render = depth_mask, see Renderer.render_all);
    r_tex: soft-texture render; r_nocs: NOCS render (channels 0..2 = match_gt, 3 = match_mask; no gradient)."""

    @staticmethod
    def forward(ctx, r_depth, r_tex, match_lr, img, mask, depth, r_nocs, hf, wf, use_depth):
        B, H, W = check_shapes(r_depth, r_tex, match_lr, img, mask, depth if use_depth else None, r_nocs, hf, wf)
        dev = r_depth.device
        r_depth, r_tex, r_nocs = r_depth.detach(), r_tex.detach(), r_nocs.detach()
        match_lr = match_lr.detach().contiguous()
        maps = [_view(img, 3, H, W), _view(mask, 1, H, W), _view(depth if use_depth else None, 1, H, W),
                _view(r_depth[:, 3], 1, H, W), _view(r_tex[:, :3], 3, H, W), _view(r_tex[:, 3], 1, H, W),
                _view(r_depth[:, 2], 1, H, W), _view(r_depth[:, 3], 1, H, W), _view(r_nocs[:, :3], 3, H, W),
                _view(r_nocs[:, 3], 1, H, W), (None, 0)]
        L = _lib.lib()
        losses = torch.empty(B, 4, dtype=torch.float32, device=dev)
        ws = torch.empty(L.scp_image_losses_workspace_bytes(B), dtype=torch.uint8, device=dev)
        ptrs, strides = _pack(maps)
        with torch.cuda.device(dev):
            rc = L.scp_image_losses_forward(ptrs, strides, _lib.ptr(match_lr), B, H, W, hf, wf, int(bool(use_depth)),
                                            _lib.ptr(losses), _lib.ptr(ws), _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_image_losses_forward')
        ctx.save_for_backward(r_depth, r_tex, match_lr, img, mask, depth if use_depth else None, r_nocs, ws)
        ctx.geom = (B, H, W, hf, wf, bool(use_depth))
        return losses

    @staticmethod
    def backward(ctx, g_losses):
        r_depth, r_tex, match_lr, img, mask, depth, r_nocs, ws = ctx.saved_tensors
        B, H, W, hf, wf, use_depth = ctx.geom
        dev = r_depth.device
        maps = [_view(img, 3, H, W), _view(mask, 1, H, W), _view(depth, 1, H, W),
                _view(r_depth[:, 3], 1, H, W), _view(r_tex[:, :3], 3, H, W), _view(r_tex[:, 3], 1, H, W),
                _view(r_depth[:, 2], 1, H, W), _view(r_depth[:, 3], 1, H, W), _view(r_nocs[:, :3], 3, H, W),
                _view(r_nocs[:, 3], 1, H, W), (None, 0)]
        g_depth = torch.zeros_like(r_depth)          # channels 0, 1 carry no gradient
        g_tex = torch.empty_like(r_tex)
        g_match = torch.empty_like(match_lr)
        gmaps = [_view(g_depth[:, 3], 1, H, W), _view(g_tex[:, :3], 3, H, W), _view(g_tex[:, 3], 1, H, W),
                 _view(g_depth[:, 2], 1, H, W) if use_depth else (None, 0), (None, 0)]
        ptrs, strides = _pack(maps)
        gptrs, gstrides = _pack(gmaps)
        g_losses = g_losses.float().contiguous()
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_image_losses_backward(ptrs, strides, _lib.ptr(match_lr), B, H, W, hf, wf, int(use_depth),
                                                      _lib.ptr(g_losses), _lib.ptr(ws), gptrs, gstrides,
                                                      _lib.ptr(g_match), _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_image_losses_backward')
        return g_depth, g_tex, g_match, None, None, None, None, None, None, None


def image_losses(r_depth, r_tex, match_lr, img, mask, depth, r_nocs, hf, wf, use_depth=True):
    """Per-image (mask_loss, texture_loss, depth_loss, match_loss), each (B,), un-weighted -- the values of
    compute_mask_loss / compute_texture_loss / compute_depth_loss[0] / compute_match_loss of loss_utils."""
    if not r_depth.is_cuda:
        raise TypeError('image_losses supports only CUDA tensors (no CPU path)')
    out = ImageLossFunction.apply(r_depth, r_tex, match_lr, img, mask, depth, r_nocs, hf, wf, use_depth)
    return out[:, 0], out[:, 1], out[:, 2], out[:, 3]
