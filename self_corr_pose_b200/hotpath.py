"""The per-image self-supervised hot path as one callable: correspondence -> texture -> SoftRas renders
-> silhouette / texture / depth / correspondence / cycle losses, forward + backward.

This is lines 73-134 of the reference's MeshNet.forward (model/model.py) with the image/mesh encoder
outputs (`img_feat`, `mesh_feat`, `pred_v`, `rotation`, `translation`) supplied by the caller.  The
full model (self_corr_pose_b200.model.model.MeshNet) calls the same stages after its encoder.

Stages are enabled as their sm_100a kernels land; a stage that is enabled always runs natively (there is
no PyTorch fallback for SoftRas / correspondence / ViT inside an enabled stage).
"""
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from .model.module.renderer import Renderer
from .model.util import loss_utils as L


def default_opts(**over):
    """Flag values of config/laptop_wild6d/base_config.txt + the flag defaults they do not override."""
    o = dict(img_size=256, batch_size=16, repeat=4, corr_h=64, corr_w=64, n_corr_feat=64, tau_img=10., tau_mesh=10.,
             use_depth=True, use_occ=False, train=True, divide_fn='both', pretrain_k=200, total_iters=20000,
             mask_wt=0.15, tex_wt=0.05, depth_wt=0.1, triangle_wt=0.002, pullfar_wt=0.01, deform_wt=0.4,
             symmetry_wt=0.5, camera_wt=0.005, match_wt=0.02, imatch_wt=0.02, decay_ratio=0.1,
             cycle_loss_wt=0.01, cycle_loss_pretrain_wt=0.02, topk_img=100, topk_mesh=100)
    o.update(over)
    return SimpleNamespace(**o)


class HotPath:
    """render stage: Renderer.render_all + the five render/correspondence losses."""

    def __init__(self, opts, mesh, stages=('softras',)):
        self.opts = opts
        self.mesh = mesh
        self.stages = tuple(stages)
        self.renderer = Renderer(opts, mesh)

    def forward(self, data, enc):
        """data = (img, mask, depth, foc_crop, pp_crop); enc = dict(pred_v, rotation, translation, tex,
        match, imatch[, img_feat, mesh_feat]).  Returns (total_loss, aux)."""
        opts = self.opts
        img, mask, depth, foc_crop, pp_crop = data
        pred_v, rotation, translation = enc['pred_v'], enc['rotation'], enc['translation']
        bsz = img.shape[0]
        faces = self.mesh.faces[None].repeat(bsz, 1, 1)
        tex, match, imatch = enc['tex'], enc['match'], enc['imatch']

        (mask_render, tex_render, depth_render, match_gt, imatch_gt, tex_mask, depth_mask, match_mask,
         depth_weight) = self.renderer.render_all(pred_v, faces, tex, foc_crop, pp_crop, rotation, translation)

        aux = {}
        aux['mask_loss'] = opts.mask_wt * L.compute_mask_loss(img, mask, mask_render).mean(0)
        aux['texture_loss'] = opts.tex_wt * L.compute_texture_loss(img, mask, tex_render, tex_mask).mean(0)
        if opts.use_depth:
            d_loss, _ = L.compute_depth_loss(depth, depth_render, depth_mask, mask)
            aux['depth_loss'] = opts.depth_wt * d_loss.mean(0)
        aux['match_loss'] = opts.match_wt * L.compute_match_loss(match, match_gt, match_mask, mask).mean(0)
        aux['imatch_loss'] = opts.imatch_wt * L.compute_imatch_loss(imatch, imatch_gt, depth_weight).mean(0)
        aux['pullfar_loss'] = opts.pullfar_wt * F.relu(1 - translation[:, :, -1]).mean()
        total = sum(aux.values())
        aux['total_loss'] = total
        aux['mask_render'] = mask_render
        aux['depth_weight'] = depth_weight
        return total, aux
