"""Trainable encoder networks (host code: PyTorch / cuDNN on both sides -- SURVEY.md section 2.1 "stays PyTorch").

Architectures and parameter names follow the reference so its checkpoints load unchanged:
  ResNetEncoder   model/module/network/image_encoder.py:119-139  (torchvision resnet18 without fc, 4 pyramid levels)
  ResNetDecoder   model/module/network/image_encoder.py:141-193  (3 x [bilinear up, conv+LeakyReLU(0.1), concat, conv], 1x1 proj)
  MeshEncoder     model/module/network/mesh_encoder.py:6-39      (PointNet-lite: STN + 1x1 conv)
  PosePredictor   model/module/network/pose_predictor.py:22-87   (6-D rotation head with fixed offsets, translation head)
  ShapePredictor  model/module/network/shape_predictor.py:12-43  (CondNeRFModel with 2 layers, third-party/nerf/models.py:336-417)
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

from .... import _flags
from ....ops import nhwc


class _CBR(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.cbr_unit = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, stride=1, bias=True),
                                      nn.LeakyReLU(0.1, inplace=True))

    def forward(self, x):
        return self.cbr_unit(x)


class ResNetEncoder(nn.Module):
    """ImageNet-pretrained ResNet-18 trunk (the reference: torchvision resnet18(pretrained=True), image_encoder.py:122).
    The weights are read from `weights_path` / $SCP_RESNET18_WEIGHTS / the torch hub cache (no download: there is no
    network); when none exists this raises like a failed download would, unless SCP_SYNTHETIC_WEIGHTS=1 (tests and
    benchmarks: seeded random initialisation)."""
    HUB_FILE = 'resnet18-f37072fd.pth'

    def __init__(self, pretrained=True, weights_path=None):
        super().__init__()
        self.resnet = torchvision.models.resnet18(weights=None)
        if pretrained:
            cands = [weights_path, os.environ.get('SCP_RESNET18_WEIGHTS'),
                     os.path.join(torch.hub.get_dir(), 'checkpoints', self.HUB_FILE)]
            path = next((c for c in cands if c and os.path.exists(c)), None)
            if path is not None:
                self.resnet.load_state_dict(torch.load(path, map_location='cpu'))
            elif not _flags.synthetic_weights_allowed():
                raise FileNotFoundError('ImageNet ResNet-18 weights not found (looked in %s); pass pretrained=False or set '
                                        'SCP_SYNTHETIC_WEIGHTS=1 for a seeded random initialisation' % [c for c in cands if c])
        self.resnet.fc = None

    def forward(self, x):
        r = self.resnet
        if x.shape[1] in (4, 8):     # zero-padded input channels (ops/color_jitter.py pad_c4): same values, cuDNN's vectorised NHWC kernels
            w = F.pad(r.conv1.weight, (0, 0, 0, 0, 0, x.shape[1] - 3)).contiguous(memory_format=torch.channels_last)
            h = F.conv2d(x, w, None, r.conv1.stride, r.conv1.padding)
        else:
            h = r.conv1(x)
        h = r.relu(r.bn1(h))
        c1 = nhwc.maxpool3x3s2(h) if nhwc.usable(h) else r.maxpool(h)     # csrc/scp_nhwc.cu on the GPU
        c2 = r.layer1(c1)
        c3 = r.layer2(c2)
        c4 = r.layer3(c3)
        c5 = r.layer4(c4)
        return c2, c3, c4, c5


class ResNetDecoder(nn.Module):
    def __init__(self, is_proj=True, out_channel=64, downsample=4):
        super().__init__()
        self.is_proj, self.downsample = is_proj, downsample
        self.upconv5, self.iconv4 = _CBR(512, 256), _CBR(512, 256)
        self.upconv4, self.iconv3 = _CBR(256, 128), _CBR(256, 128)
        self.upconv3, self.iconv2 = _CBR(128, 64), _CBR(128, 64)
        if is_proj:
            self.proj = nn.Conv2d(64 if downsample == 4 else 128, out_channel, 1)

    def forward(self, c2, c3, c4, c5):
        def up(x, ref):
            if nhwc.usable(x) and ref.shape[2] == 2 * x.shape[2] and ref.shape[3] == 2 * x.shape[3]:
                return nhwc.upsample2x(x)                                  # csrc/scp_nhwc.cu on the GPU
            return F.interpolate(x, ref.shape[2:], mode='bilinear', align_corners=False)
        c4 = self.iconv4(torch.cat((c4, self.upconv5(up(c5, c4))), dim=1))
        c3 = self.iconv3(torch.cat((c3, self.upconv4(up(c4, c3))), dim=1))
        c2 = self.iconv2(torch.cat((c2, self.upconv3(up(c3, c2))), dim=1))
        out = c2 if self.downsample == 4 else c3
        return self.proj(out) if self.is_proj else out


class _STN(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv1d(3, 128, 1)
        self.fc = nn.Linear(128, 9)

    def forward(self, x):                      # b,3,n -> b,3,3 (identity + residual)
        h = F.relu(self.conv1(x)).max(2)[0]
        return (self.fc(h) + torch.eye(3, device=x.device).reshape(1, 9)).view(-1, 3, 3)


class MeshEncoder(nn.Module):
    def __init__(self, n_feat):
        super().__init__()
        self.stn = _STN()
        self.conv1 = nn.Conv1d(3, n_feat, 1)

    def forward(self, x):                      # b,n,3 -> b,n,c
        trans = self.stn(x.transpose(2, 1))
        y = torch.bmm(x, trans).transpose(2, 1)
        return F.relu(self.conv1(y)).transpose(2, 1)


def _fc_stack(nc_inp, nc_out, nlayers):
    mods = []
    for _ in range(nlayers):
        mods.append(nn.Sequential(nn.Linear(nc_inp, nc_out), nn.LeakyReLU(0.1, inplace=True)))
        nc_inp = nc_out
    net = nn.Sequential(*mods)
    for m in net.modules():
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(0, 0.02)
            m.bias.data.zero_()
    return net


class PosePredictor(nn.Module):
    def __init__(self, opts, nc_input):
        super().__init__()
        self.offset = opts.depth_offset
        self.use_scale = getattr(opts, 'use_scale', False)
        self.n_hypo = 1                         # the reference asserts n_hypo == 1 (pose_predictor.py:32)
        self.rot_pred_layer = nn.Sequential(_fc_stack(nc_input, 128, 3), nn.Linear(128, 6))
        self.trans_pred_layer = nn.Linear(nc_input, 3)
        if self.use_scale:
            self.scale_pred_layer = nn.Linear(nc_input, 3)
        r = [float(v) for v in opts.rotation_offset]
        self.x_offset = nn.Parameter(torch.tensor([r[:3]]), requires_grad=False)
        self.y_offset = nn.Parameter(torch.tensor([r[3:]]), requires_grad=False)

    def forward(self, feat):
        bsz = feat.shape[0]
        rot = self.rot_pred_layer(feat).reshape(bsz, 6)
        trans = self.trans_pred_layer(feat).reshape(bsz, 3)
        x = F.normalize(rot[:, :3] + self.x_offset)
        y = rot[:, 3:6] + self.y_offset
        z = F.normalize(torch.cross(x, y, dim=-1))
        y = F.normalize(torch.cross(z, x, dim=-1))
        rot = torch.stack((x, y, z), 2)         # Gram-Schmidt of the 6-D representation, columns x,y,z
        trans = torch.cat((trans[:, :2] * 0.1, trans[:, 2:] + self.offset), dim=1)
        if self.use_scale:
            scale = self.scale_pred_layer(feat).reshape(bsz, 3) * 0.1 + 1.
        else:
            scale = torch.ones((bsz, 3), device=feat.device)
        return rot, trans, scale


class CondMLP(nn.Module):
    """CondNeRFModel(num_layers=2, no positional encoding, no view dirs): xyz ++ code -> 256 -> 256 -> (rgb 3, alpha 1)."""

    def __init__(self, codesize, hidden=256, out_channel=3):
        super().__init__()
        self.codesize = codesize
        self.layer1 = nn.Linear(3 + codesize, hidden)
        self.layers_xyz = nn.ModuleList([nn.Linear(hidden, hidden)])
        self.layers_dir = nn.ModuleList([nn.Linear(hidden, hidden // 2)])
        self.fc_alpha = nn.Linear(hidden, 1)
        self.fc_rgb = nn.Linear(hidden // 2, out_channel)
        self.fc_feat = nn.Linear(hidden, hidden)

    def forward(self, x):
        h = F.relu(self.layers_xyz[0](self.layer1(x)))
        alpha = self.fc_alpha(h)
        rgb = self.fc_rgb(F.relu(self.layers_dir[0](F.relu(self.fc_feat(h)))))
        return torch.cat((rgb, alpha), dim=-1)


class ShapePredictor(nn.Module):
    def __init__(self, opts):
        super().__init__()
        self.shapenerf = CondMLP(opts.codedim)
        self.no_deform = getattr(opts, 'no_deform', False)
        self.deform_ratio = getattr(opts, 'deform_ratio', 1.)

    def forward(self, mean_v, shape_code):
        if self.no_deform:
            return mean_v
        b, n = mean_v.shape[:2]
        inp = torch.cat([mean_v.detach().reshape(-1, 3), shape_code[:, None].repeat(1, n, 1).view(-1, shape_code.shape[-1])], 1)
        delta = self.shapenerf(inp).reshape(b, n, -1)[:, :, :-1]
        delta = delta - delta.mean(1, keepdim=True)
        return mean_v + delta * self.deform_ratio
