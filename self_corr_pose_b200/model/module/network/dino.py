"""Frozen DINO ViT-S/8 feature extractor on the sm_100a tcgen05 kernels.

API of the reference's model/module/network/dino.py (`DINO()`, `forward(img) -> (b, 384, h/8, w/8)`), with
`self.model` holding the parameters under the reference checkpoint's names so that
`pretrain_corr_net.net.model.*` keys load unchanged.  Like the reference (network/dino.py:42: torch.load of
`pretrain/dino_deitsmall8_pretrain.pth`), a missing checkpoint is an error; seeded synthetic weights (vit_weights.py) are
used only when a state dict is passed or the caller opts in with SCP_SYNTHETIC_WEIGHTS=1 (_flags.py).
"""
import ctypes
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import _flags, _lib
from . import vit_weights as VW


class ViTSmall8Params(nn.Module):
    """Parameter container with the names of vision_transformer_flexible.VisionTransformer (vit_small, patch 8)."""

    def __init__(self, state_dict=None):
        super().__init__()
        sd = state_dict if state_dict is not None else VW.synthetic_state_dict(0)
        for name, shape in VW.vit_small_shapes().items():
            t = sd[name].detach().clone().float()
            assert tuple(t.shape) == tuple(shape), (name, t.shape, shape)
            self._register(name, t)

    def _register(self, dotted, tensor):
        mod = self
        parts = dotted.split('.')
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def resampled_pos_embed(pos_embed, w_patches, h_patches):
    """interpolate_pos_encoding (vision_transformer_flexible.py:192-212): bicubic with scale (n+0.1)/n0."""
    n0 = pos_embed.shape[1] - 1
    if w_patches * h_patches == n0 and w_patches == h_patches:
        return pos_embed[0]
    dim = pos_embed.shape[-1]
    s0 = int(math.sqrt(n0))
    patch = F.interpolate(pos_embed[:, 1:].reshape(1, s0, s0, dim).permute(0, 3, 1, 2).float().cpu(),
                          scale_factor=((w_patches + 0.1) / s0, (h_patches + 0.1) / s0), mode='bicubic')
    assert patch.shape[-2] == w_patches and patch.shape[-1] == h_patches
    patch = patch.permute(0, 2, 3, 1).reshape(-1, dim).to(pos_embed.device)
    return torch.cat((pos_embed[0, :1], patch), dim=0)


def split_bf16_i32(t):
    """fp32 [R, K] -> bf16 [R, 2K] split pairs in the "i32" layout of csrc/scp_gemm.cuh: hi = bf16(x), lo = bf16(x - hi),
    groups of 32 columns stored as [32 hi | 32 lo] -- the operand format of the fp32-class (x3) tensor-core products."""
    t = t.float()
    R, K = t.shape
    assert K % 32 == 0
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.stack((hi.reshape(R, K // 32, 32), lo.reshape(R, K // 32, 32)), dim=2).reshape(R, 2 * K).contiguous()


def merge_bf16_i32(t):
    """Inverse of split_bf16_i32 up to the split's own rounding: bf16 [..., 2K] -> fp32 [..., K] = hi + lo."""
    sh = t.shape
    g = t.float().reshape(*sh[:-1], sh[-1] // 64, 2, 32)
    return (g[..., 0, :] + g[..., 1, :]).reshape(*sh[:-1], sh[-1] // 2)


class DINO(nn.Module):
    """precision: 'x3' (default; fp32-class split-bf16 products, the mode that meets the parity contract against the
    reference's fp32 ViT) or 'bf16' (labelled fast mode, features 5e-3 off); env SCP_VIT_PRECISION overrides the default."""
    feat_layer = 9
    patch_size = 8
    pretrain_path = 'pretrain/dino_deitsmall8_pretrain.pth'

    def __init__(self, state_dict=None, precision=None):
        super().__init__()
        precision = precision or os.environ.get('SCP_VIT_PRECISION', 'x3')
        if precision not in ('x3', 'bf16'):
            raise ValueError("DINO precision must be 'x3' or 'bf16', got %r" % (precision,))
        self.precision = precision
        if state_dict is None:
            if os.path.exists(self.pretrain_path):
                state_dict = torch.load(self.pretrain_path, map_location='cpu')
            elif not _flags.synthetic_weights_allowed():
                raise FileNotFoundError(
                    'DINO checkpoint %s not found (the pseudo-ground-truth matches are meaningless without it); set '
                    'SCP_SYNTHETIC_WEIGHTS=1 to run with seeded synthetic weights (tests / benchmarks only)' % self.pretrain_path)
        self.model = ViTSmall8Params(state_dict)
        self._packed = {}

    def _apply(self, fn, *a, **k):
        self._packed = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = {}
        return super().load_state_dict(*a, **k)

    def _pack(self, H, W, device):
        """bf16 GEMM weights, fp32 vectors and the resampled position embedding, as the C-ABI struct."""
        key = (H, W, str(device), self.precision)
        x3 = self.precision == 'x3'
        if key in self._packed:
            return self._packed[key]
        sd = {k: v.detach().to(device) for k, v in self.model.state_dict().items()}
        keep = []

        def f32(t):
            t = t.float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def b16(t):
            t = split_bf16_i32(t) if x3 else t.to(torch.bfloat16).contiguous()
            keep.append(t)
            return t.data_ptr()

        pos = resampled_pos_embed(sd['pos_embed'], W // 8, H // 8)
        w = _lib.VitWeights()
        w.patch_w = b16(sd['patch_embed.proj.weight'].reshape(VW.EMBED, -1))
        w.patch_b = f32(sd['patch_embed.proj.bias'])
        w.cls_pos0 = f32(sd['cls_token'].reshape(-1) + pos[0])
        w.pos = f32(pos[1:])
        for i in range(VW.DEPTH):
            p = 'blocks.%d.' % i
            blk = w.blocks[i]
            blk.ln1_w, blk.ln1_b = f32(sd[p + 'norm1.weight']), f32(sd[p + 'norm1.bias'])
            blk.qkv_w, blk.qkv_b = b16(sd[p + 'attn.qkv.weight']), f32(sd[p + 'attn.qkv.bias'])
            blk.proj_w, blk.proj_b = b16(sd[p + 'attn.proj.weight']), f32(sd[p + 'attn.proj.bias'])
            blk.ln2_w, blk.ln2_b = f32(sd[p + 'norm2.weight']), f32(sd[p + 'norm2.bias'])
            blk.fc1_w, blk.fc1_b = b16(sd[p + 'mlp.fc1.weight']), f32(sd[p + 'mlp.fc1.bias'])
            blk.fc2_w, blk.fc2_b = b16(sd[p + 'mlp.fc2.weight']), f32(sd[p + 'mlp.fc2.bias'])
        self._packed[key] = (w, keep)
        return self._packed[key]

    @torch.no_grad()
    def forward(self, img, layer=None, tokens=False):
        """img (b,3,H,W) raw [0,1] RGB -> layer-9 key features (b, 384, H/8, W/8); frozen, no autograd.
        tokens=True additionally returns the same features token-major in bf16, the operand of the native arg-max
        matching (PretrainedCorrespondence): (b, H/8*W/8, 384), or split pairs (b, H/8*W/8, 768) in x3 precision."""
        if not img.is_cuda:
            raise TypeError('DINO supports only CUDA tensors (no CPU path)')
        layer = self.feat_layer if layer is None else layer
        B, _, H, W = img.shape
        dev = img.device
        w, _keep = self._pack(H, W, dev)
        img = img.detach().float().contiguous()
        feat = torch.empty(B, VW.EMBED, H // 8, W // 8, dtype=torch.float32, device=dev)
        L = _lib.lib()
        prec = _lib.VIT_X3 if self.precision == 'x3' else _lib.VIT_BF16
        ws_bytes = L.scp_vit_workspace_bytes(B, H, W, prec)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        tok_w = VW.EMBED * (2 if prec == _lib.VIT_X3 else 1)
        tok = torch.empty(B, (H // 8) * (W // 8), tok_w, dtype=torch.bfloat16, device=dev) if tokens else None
        with torch.cuda.device(dev):
            rc = L.scp_vit_s8_keys(ctypes.byref(w), _lib.ptr(img), _lib.ptr(feat), _lib.ptr(tok), B, H, W, layer, prec,
                                   _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_vit_s8_keys')
        return (feat, tok) if tokens else feat
