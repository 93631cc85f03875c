"""The hot path on the CPU in the reference's own formulation -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Used by tests (step-level parity of HotPath), __graft_entry__.smoke() and bench.py's `cpu_baseline` /
`--impl reference` legs.  Every stage follows the reference:
  correspondence  : oracle/corr.py::match          (materialised (B,P,N) softmaxes, correspondence.py:36-73)
  SoftRas         : oracle/softras_oracle.c         (one pixel x ALL faces, soft_rasterize_cuda_kernel.cu)
                    or oracle/_ref (the reference's kernel source built for the host) when use_ref=True
  DINO ViT        : oracle/vit.py on cat(src, tgt) of every pair (4B images, like pretrained_corr.py:57-71)
  pre-train cycle : oracle/corr.py::pretrain_match / pretrain_cycle_loss (the (2B,1024,1024) corr is formed)
  camera, losses  : the device-agnostic torch code of self_corr_pose_b200.model (same statements as
                    model/util/loss_utils.py), with the rasteriser call swapped for the CPU one.
"""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F
from torch.autograd import Function

from oracle import corr as ocorr
from oracle import softras as osr
from oracle import vit as ovit


class _CpuSoftRasterize(Function):
    @staticmethod
    def forward(ctx, face_vertices, textures, kw, use_ref, nthreads, fma=False):
        fwd = osr.ref_forward if use_ref else osr.forward
        extra = {} if use_ref else {'nthreads': nthreads, 'fma': fma}
        col, info, aggr = fwd(face_vertices.detach().numpy(), textures.detach().numpy(), **kw, **extra)
        ctx.kw, ctx.use_ref, ctx.nthreads, ctx.fma = kw, use_ref, nthreads, fma
        ctx.save_for_backward(face_vertices.detach(), textures.detach())
        ctx.np_saved = (col, info, aggr)
        return torch.from_numpy(col)

    @staticmethod
    def backward(ctx, g):
        fv, tex = ctx.saved_tensors
        col, info, aggr = ctx.np_saved
        bwd = osr.ref_backward if ctx.use_ref else osr.backward
        extra = {} if ctx.use_ref else {'nthreads': ctx.nthreads, 'fma': ctx.fma}
        gf, gt = bwd(fv.numpy(), tex.numpy(), col, info, aggr, g.contiguous().numpy(), **ctx.kw, **extra)
        return torch.from_numpy(gf).reshape(fv.shape), torch.from_numpy(gt).reshape(tex.shape), None, None, None, None


@contextlib.contextmanager
def cpu_rasterizer(use_ref=False, nthreads=0, fma=False):
    """Routes soft_renderer.functional.soft_rasterize to the CPU checker while the context is active
    (oracle harness only; the product never does this)."""
    from self_corr_pose_b200.soft_renderer import functional as srf

    def soft_rasterize(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                       fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4,
                       aggr_func_rgb='softmax', aggr_func_alpha='prod', texture_type='surface'):
        kw = dict(image_size=image_size, background_color=tuple(background_color), near=near, far=far,
                  fill_back=fill_back, eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps,
                  gamma_val=gamma_val, aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha,
                  texture_type=texture_type)
        return _CpuSoftRasterize.apply(face_vertices.float().contiguous(), textures.float().contiguous(), kw, use_ref,
                                       nthreads, fma)
    saved = srf.soft_rasterize
    srf.soft_rasterize = soft_rasterize
    try:
        yield
    finally:
        srf.soft_rasterize = saved


def make_batch_cpu(opts, verts, faces, B, seed=0, nthreads=0):
    """synthetic.make_batch on the CPU (ground truth rasterised by the oracle)."""
    from types import SimpleNamespace
    from self_corr_pose_b200 import synthetic
    from self_corr_pose_b200.model.module.renderer import Renderer
    mesh = SimpleNamespace(mean_v=torch.from_numpy(verts), faces=torch.from_numpy(faces), texture_type='vertex')
    with cpu_rasterizer(nthreads=nthreads):
        return synthetic.make_batch(opts, verts, faces, B, device='cpu', seed=seed, renderer=Renderer(opts, mesh))


def forward(opts, mean_v, faces, data, enc, vit_sd, it=0, use_ref=False, nthreads=0, all_vit_blocks=False, fma=False):
    """Reference-formulation forward of HotPath.forward on CPU tensors.  Returns (total, aux)."""
    from types import SimpleNamespace
    from self_corr_pose_b200.model.module.renderer import Renderer
    from self_corr_pose_b200.model.module.weights import Weights
    from self_corr_pose_b200.model.util import loss_utils as L
    wts = Weights(opts)
    wts.schedule(it)
    img, mask, depth, foc_crop, pp_crop = data
    img_feat, mesh_feat, pred_v, rotation, translation = enc
    bsz = img.shape[0]
    hf, wf = opts.corr_h, opts.corr_w
    mesh = SimpleNamespace(mean_v=mean_v, faces=faces, texture_type='vertex')
    fb = faces[None].repeat(bsz, 1, 1)
    mv = mean_v[None].repeat(bsz, 1, 1)

    pointcorr, match, imatch, _ = ocorr.match(img_feat, mesh_feat, mask, pred_v, hf, wf, opts.tau_img, opts.tau_mesh)
    tex = F.grid_sample(img, imatch.permute(0, 2, 1)[:, None], align_corners=False)[:, :, 0].permute(0, 2, 1)
    with cpu_rasterizer(use_ref=use_ref, nthreads=nthreads, fma=fma):
        (mask_render, tex_render, depth_render, match_gt, imatch_gt, tex_mask, depth_mask, match_mask,
         depth_weight) = Renderer(opts, mesh, reference_launches=True).render_all(pred_v, fb, tex, foc_crop, pp_crop, rotation,
                                                                                  translation)
    aux = {}
    aux['mask_loss'] = wts.mask_wt * L.compute_mask_loss(img, mask, mask_render).mean(0)
    aux['texture_loss'] = wts.tex_wt * L.compute_texture_loss(img, mask, tex_render, tex_mask).mean(0)
    if opts.use_depth:
        d_loss, _ = L.compute_depth_loss(depth, depth_render, depth_mask, mask)
        aux['depth_loss'] = wts.depth_wt * d_loss.mean(0)
    aux['match_loss'] = wts.match_wt * L.compute_match_loss(match, match_gt, match_mask, mask).mean(0)
    aux['imatch_loss'] = wts.imatch_wt * L.compute_imatch_loss(imatch, imatch_gt, depth_weight).mean(0)
    aux['triangle_loss'] = wts.triangle_wt * L.LaplacianLoss(mean_v, faces, average=True)(pred_v) * pred_v.shape[1] / 64.
    aux['pullfar_loss'] = wts.pullfar_wt * F.relu(1 - translation[:, :, -1]).mean()
    aux['deform_loss'] = wts.deform_wt * F.smooth_l1_loss(pred_v, mv, reduction='mean')

    # pre-training cycle loss, reference formulation (pretrained_corr.py:107-140)
    fn = ocorr.DIVIDE[opts.divide_fn]
    bs, rep = opts.batch_size, opts.repeat
    img_src, img_tgt = fn(img, bs, rep)
    mask_src, mask_tgt = fn(mask, bs, rep)
    nb = img_src.shape[0]
    grid = ocorr.meshgrid(hf, wf).reshape(2, hf, wf)[None].repeat(nb, 1, 1, 1)
    grid = F.interpolate(grid, (hf // 2, wf // 2), mode='bilinear')
    with torch.no_grad():
        layer = 11 if all_vit_blocks else 9
        if all_vit_blocks:   # run the three discarded blocks too, like the reference does
            ovit.keys_at_layer(vit_sd, torch.cat([img_src, img_tgt])[:1], layer, 6)
        feat = ovit.dino_features(vit_sd, torch.cat([img_src, img_tgt], 0))
        pts_src, pts_tgt, i_src, i_tgt, mask_k = ocorr.pretrain_match(feat[:nb], feat[nb:], mask_src, mask_tgt, grid,
                                                                    opts.img_size // 8, opts.pretrain_k)
    cyc, _ = ocorr.pretrain_cycle_loss(pts_src, i_tgt, mask_k, depth_weight, pointcorr, bs, rep, hf, wf,
                                       opts.divide_fn, opts.tau_img, opts.tau_mesh)
    aux['cycle_loss_pretrain'] = cyc * wts.cycle_loss_pt_wt
    total = sum(aux.values())
    aux['total_loss'] = total
    aux['mask_render'] = mask_render
    aux['match'] = match
    aux['imatch'] = imatch
    aux['_pointcorr'], aux['_depth_weight'] = pointcorr, depth_weight      # for oracle/trainer_cpu.py
    return total, aux


def step(opts, mean_v, faces, data, enc, vit_sd, **kw):
    for t in enc:
        t.grad = None
    total, aux = forward(opts, mean_v, faces, data, enc, vit_sd, **kw)
    total.backward()
    return total, aux
