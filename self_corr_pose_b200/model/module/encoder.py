"""Encoder: image -> (image code, per-pixel features), mesh -> per-vertex features, pose and shape heads.
API and statements of the reference's model/module/encoder.py:13-52 (host code, PyTorch)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

from .network.encoder_nets import ResNetEncoder, ResNetDecoder, MeshEncoder, PosePredictor, ShapePredictor


class Encoder(nn.Module):

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.resnet_transform = torchvision.transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        self.random_jitter = torchvision.transforms.ColorJitter(0.2, 0.2, 0.2, 0.05)
        self.backbone = ResNetEncoder()
        self.featnet = ResNetDecoder(is_proj=True, out_channel=opts.n_corr_feat, downsample=opts.img_size // opts.corr_h)
        self.featnet_mesh = MeshEncoder(opts.n_corr_feat)
        self.shape_code_predictor = nn.Linear(512, opts.codedim)
        self.shape_predictor = ShapePredictor(opts)
        self.pose_predictor = PosePredictor(opts, 512)

    def encode_img(self, img):
        bsz = img.shape[0]
        x = self.resnet_transform(self.random_jitter(img))       # jitter runs in eval too (encoder.py:31)
        c2, c3, c4, c5 = self.backbone(x)
        img_code = c5.mean((2, 3))
        img_feat = self.featnet(c2, c3, c4, c5).reshape(bsz, self.opts.n_corr_feat, -1)
        return img_code, F.normalize(img_feat, 2, 1)

    def forward(self, img, mean_v, pp_crop, foc_crop):
        img_code, img_feat = self.encode_img(img)
        pred_v = self.shape_predictor(mean_v, self.shape_code_predictor(img_code))
        mesh_feat = F.normalize(self.featnet_mesh(pred_v.detach()), 2, -1)
        rotation, translation, scale = self.pose_predictor(img_code)
        pred_v = pred_v * scale[:, None]
        # principal-point shift with detached depth (encoder.py:49); out-of-place form of the reference's `-=`
        shift = (pp_crop / foc_crop) * translation[:, 2:].detach()
        translation = torch.cat(((translation[:, :2].to(shift.dtype) - shift).to(translation.dtype), translation[:, 2:]), dim=1)
        return img_feat, mesh_feat, pred_v, rotation.reshape(-1, 3, 3), translation.reshape(-1, 1, 3), scale
