"""Fused ColorJitter + Normalize of the encoder input (csrc/scp_jitter.cu).  The random parameters are drawn by
torchvision's own ColorJitter.get_params (same consumption of the global CPU generator as the reference's
`self.random_jitter(img)`, model/module/encoder.py:31); only the application is native.  No autograd: the encoder input is
data.  No CPU path."""
import ctypes

import numpy as np
import torch
from torchvision import transforms

from .. import _lib


def jitter_normalize(img, jitter, mean, std, params=None, channels_last=True):
    """img (B,3,H,W) fp32 CUDA in [0,1]; jitter: a torchvision ColorJitter; returns Normalize(mean,std)(jitter(img)).
    params = (fn_idx, b, c, s, h) as returned by ColorJitter.get_params, drawn here when None.
    channels_last: the result is laid out NHWC in memory (torch.channels_last strides, same logical (B,3,H,W) tensor), so
    that the ResNet that consumes it runs cuDNN's NHWC kernels end to end without layout conversions."""
    if not img.is_cuda:
        raise TypeError('jitter_normalize supports only CUDA tensors (no CPU path)')
    if params is None:
        params = transforms.ColorJitter.get_params(jitter.brightness, jitter.contrast, jitter.saturation, jitter.hue)
    fn_idx, b, c, s, h = params
    order, ratios = [], []
    for fid in [int(i) for i in fn_idx]:
        f = (b, c, s, h)[fid]
        order.append(fid if f is not None else -1)
    for f in (b, c, s):
        f = 1.0 if f is None else float(f)
        ratios += [np.float32(f), np.float32(1.0 - f)]      # _blend: `ratio * img1 + (1.0 - ratio) * img2`, python doubles
    img = img.detach().float().contiguous()
    B, _, H, W = img.shape
    out = torch.empty_like(img, memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    L = _lib.lib()
    dev = img.device
    ws_bytes = L.scp_color_jitter_workspace_bytes(B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    c_order = (ctypes.c_int * 4)(*order)
    c_rat = (ctypes.c_float * 6)(*[float(r) for r in ratios])
    c_mean = (ctypes.c_float * 3)(*[float(np.float32(m)) for m in mean])
    c_std = (ctypes.c_float * 3)(*[float(np.float32(m)) for m in std])
    with torch.cuda.device(dev):
        rc = L.scp_color_jitter_normalize(_lib.ptr(img), out.data_ptr(), B, H * W, c_order, c_rat,
                                          0.0 if h is None else float(h), c_mean, c_std, 1 if channels_last else 0,
                                          _lib.ptr(ws), ws_bytes,
                                          _lib.stream_ptr(dev))
    _lib.check(rc, 'scp_color_jitter_normalize')
    return out
