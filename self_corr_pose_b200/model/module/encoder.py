"""Trainable encoder of the model: ResNet-18 image branch (global code + per-pixel correspondence features), mesh
branch (per-vertex features), shape and pose heads.

Host code in PyTorch / cuDNN (the next component after the hot path), behaviour of the reference's
model/module/encoder.py:13-52 -- pinned to it by tests/test_reference_encoder_cpu.py (same parameter names, same RNG
consumption, same outputs).  Submodule names are part of the contract: checkpoints and the optimiser's name-keyed
parameter groups (model/module/optimizers.py) depend on them.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision import transforms

from .network import encoder_nets as nets
from ...ops import nhwc
from ...ops.color_jitter import JitterSlot, jitter_normalize

_IMAGENET_MEAN, _IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


class Encoder(nn.Module):

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        C = opts.n_corr_feat
        for name, module in (
                ('backbone', nets.ResNetEncoder()),
                ('featnet', nets.ResNetDecoder(is_proj=True, out_channel=C, downsample=opts.img_size // opts.corr_h)),
                ('featnet_mesh', nets.MeshEncoder(C)),
                ('shape_code_predictor', nn.Linear(512, opts.codedim)),
                ('shape_predictor', nets.ShapePredictor(opts)),
                ('pose_predictor', nets.PosePredictor(opts, 512))):
            self.add_module(name, module)
        # photometric jitter is applied in evaluation mode as well (appendix A.7 of SURVEY.md), then ImageNet statistics
        self.random_jitter = transforms.ColorJitter(0.2, 0.2, 0.2, 0.05)
        self.resnet_transform = transforms.Normalize(mean=list(_IMAGENET_MEAN), std=list(_IMAGENET_STD))

    def enable_static_params(self, device):
        """CUDA-graph mode: the jitter parameters of the two encoder passes of a step travel through device memory."""
        self.jitter_slots = [JitterSlot(device, _IMAGENET_MEAN, _IMAGENET_STD) for _ in range(2)]

    def encode_img(self, img, pass_idx=0):
        """(b,3,H,W) in [0,1] -> (global code (b,512), unit-norm pixel features (b,C,h*w)).  pass_idx: 0 = the step's first
        encoder pass, 1 = the rotated second pass (selects the static parameter slot in CUDA-graph mode)."""
        if img.is_cuda:     # one native pass instead of torchvision's ~70 launches (same random draws, ops/color_jitter.py).
            # SCP_STEM_C4=1: zero fourth input channel for the stem convolution -- measured, no gain (39.73 vs 39.65 ms per step)
            slots = getattr(self, 'jitter_slots', None)
            x = jitter_normalize(img, self.random_jitter, _IMAGENET_MEAN, _IMAGENET_STD,
                                 slot=slots[pass_idx] if slots else None,
                                 pad_c4={'1': True, '8': 8}.get(os.environ.get('SCP_STEM_C4', '0'), False))
        else:
            x = self.resnet_transform(self.random_jitter(img))
        pyramid = self.backbone(x)
        code = pyramid[-1].mean(dim=(2, 3))
        feat = self.featnet(*pyramid)
        if nhwc.usable(feat) and feat.shape[1] <= 128:       # NHWC in, contiguous (b, C, h*w) unit vectors out (csrc/scp_nhwc.cu)
            return code, nhwc.l2norm_cp(feat)
        return code, F.normalize(feat.flatten(2), p=2, dim=1)

    def _shape(self, code, mean_v):
        return self.shape_predictor(mean_v, self.shape_code_predictor(code))

    @staticmethod
    def _recentre(translation, pp_crop, foc_crop):
        """Principal-point shift of the xy translation with the depth detached.  The reference does this in place on an
        fp32 tensor with fp64 intrinsics (encoder.py:49): the update is evaluated in fp64 and rounded back."""
        z = translation[:, 2:]
        xy = translation[:, :2] - (pp_crop / foc_crop) * z.detach()
        return torch.cat((xy.to(translation.dtype), z), dim=1)

    def forward(self, img, mean_v, pp_crop, foc_crop):
        code, img_feat = self.encode_img(img)
        verts = self._shape(code, mean_v)
        mesh_feat = F.normalize(self.featnet_mesh(verts.detach()), p=2, dim=-1)
        rotation, translation, scale = self.pose_predictor(code)
        translation = self._recentre(translation, pp_crop, foc_crop)
        return (img_feat, mesh_feat, verts * scale[:, None], rotation.reshape(-1, 3, 3), translation.reshape(-1, 1, 3),
                scale)
