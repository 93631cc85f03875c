"""Fused ColorJitter + Normalize of the encoder input (csrc/scp_jitter.cu).  The random parameters are drawn by
torchvision's own ColorJitter.get_params (same consumption of the global CPU generator as the reference's
`self.random_jitter(img)`, model/module/encoder.py:31); only the application is native.  No autograd: the encoder input is
data.  No CPU path."""
import ctypes
import struct

import numpy as np
import torch
from torchvision import transforms

from .. import _flags, _lib


def _pack(params, mean, std):
    """(fn_idx, b, c, s, h) -> (order[4], ratios[6], hue): what the kernel consumes."""
    fn_idx, b, c, s, h = params
    order, ratios = [], []
    for fid in [int(i) for i in fn_idx]:
        order.append(fid if (b, c, s, h)[fid] is not None else -1)
    for f in (b, c, s):
        f = 1.0 if f is None else float(f)
        ratios += [float(np.float32(f)), float(np.float32(1.0 - f))]   # _blend: `ratio * img1 + (1.0 - ratio) * img2`, python doubles
    return order, ratios, 0.0 if h is None else float(h)


class JitterSlot:
    """Parameter block of one jitter call site in pinned host memory + its device copy (scp_jitter_params).  Inside a CUDA
    graph the site launches [copy host -> device, kernel reading the device block]; before every replay `refresh` draws new
    parameters (same torchvision call, same consumption of the CPU generator) into the host block."""
    FMT = '4i13f'

    def __init__(self, device, mean, std):
        n = struct.calcsize(self.FMT)
        self.host = _flags.pinned(torch.zeros(n, dtype=torch.uint8))
        self.dev = torch.zeros(n, dtype=torch.uint8, device=device)
        self.mean, self.std = [float(np.float32(m)) for m in mean], [float(np.float32(m)) for m in std]

    def refresh(self, jitter):
        params = transforms.ColorJitter.get_params(jitter.brightness, jitter.contrast, jitter.saturation, jitter.hue)
        order, ratios, hue = _pack(params, self.mean, self.std)
        r1, r2 = ratios[0::2], ratios[1::2]
        blob = struct.pack(self.FMT, *order, *r1, *r2, hue, *self.mean, *self.std)
        self.host.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))

    def upload(self):
        self.dev.copy_(self.host, non_blocking=True)


def jitter_normalize(img, jitter, mean, std, params=None, channels_last=True, slot=None, pad_c4=False):
    """img (B,3,H,W) fp32 CUDA in [0,1]; jitter: a torchvision ColorJitter; returns Normalize(mean,std)(jitter(img)).
    params = (fn_idx, b, c, s, h) as returned by ColorJitter.get_params, drawn here when None.
    channels_last: the result is laid out NHWC in memory (torch.channels_last strides, same logical (B,3,H,W) tensor), so
    that the ResNet that consumes it runs cuDNN's NHWC kernels end to end without layout conversions.
    slot: a JitterSlot -> the parameters travel through device memory (CUDA-graph capturable call site).
    pad_c4: channels-last result with a zero FOURTH channel, (B,4,H,W) (pad_c4 = 8: padded to eight channels) -- feed it to the stem convolution with the weight
    zero-padded to 4 input channels (same values; cuDNN's 3-channel NHWC path is a slow legacy kernel)."""
    if not img.is_cuda:
        raise TypeError('jitter_normalize supports only CUDA tensors (no CPU path)')
    img = img.detach().float().contiguous()
    B, _, H, W = img.shape
    if pad_c4:
        assert channels_last
        out = torch.empty(B, 8 if pad_c4 == 8 else 4, H, W, device=img.device, dtype=img.dtype, memory_format=torch.channels_last)
    else:
        out = torch.empty_like(img, memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    layout = (3 if pad_c4 == 8 else 2) if pad_c4 else (1 if channels_last else 0)
    L = _lib.lib()
    dev = img.device
    ws_bytes = L.scp_color_jitter_workspace_bytes(B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if slot is not None:
        if params is not None:
            raise ValueError('a static jitter slot draws its own parameters')
        slot.refresh(jitter)
        slot.upload()
        with torch.cuda.device(dev):
            rc = L.scp_color_jitter_normalize_dparams(_lib.ptr(img), out.data_ptr(), B, H * W, _lib.ptr(slot.dev),
                                                      layout, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_color_jitter_normalize_dparams')
        return out
    if params is None:
        params = transforms.ColorJitter.get_params(jitter.brightness, jitter.contrast, jitter.saturation, jitter.hue)
    order, ratios, hue = _pack(params, mean, std)
    c_order = (ctypes.c_int * 4)(*order)
    c_rat = (ctypes.c_float * 6)(*ratios)
    c_mean = (ctypes.c_float * 3)(*[float(np.float32(m)) for m in mean])
    c_std = (ctypes.c_float * 3)(*[float(np.float32(m)) for m in std])
    with torch.cuda.device(dev):
        rc = L.scp_color_jitter_normalize(_lib.ptr(img), out.data_ptr(), B, H * W, c_order, c_rat, hue, c_mean, c_std,
                                          layout, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
    _lib.check(rc, 'scp_color_jitter_normalize')
    return out
