"""CPU restatement (plain torch ops, reference formulation with every (B,P,N) tensor materialised) of the
correspondence stages -- TEST INFRASTRUCTURE ONLY.

Follows, statement by statement:
  match()                      model/module/correspondence.py:36-73 (training path; eval-only match_conf :57-69)
  rotation_cycle()             model/module/correspondence.py:76-113 with the second encoder pass and the
                               torchvision rotations supplied by the caller (they are library code on both sides)
  pretrain_match()             model/module/pretrained_corr.py:73-104 (given the DINO features)
  pretrain_cycle_loss()        model/module/pretrained_corr.py:107-140
  divide_by_*                  model/util/loss_utils.py:326-345
Pinned against the reference modules themselves (imported with stub dependencies) by
tests/golden/make_corr_golden.py -> tests/golden/corr_golden.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F


def select_topk(score, k):
    """torch.topk(-distance, k) of pretrained_corr.py:97; a hook because the choice among exactly tied distances is
    implementation-defined (see the product's model/module/pretrained_corr.py::select_topk)."""
    return torch.topk(score, k=k, dim=1).indices


def stable_topk(score, k):
    """Deterministic variant for CPU <-> GPU parity tests: descending score, lowest index first among equal scores."""
    return torch.sort(score, dim=1, descending=True, stable=True).indices[:, :k]


def meshgrid(hf, wf):
    g = torch.Tensor(np.array(np.meshgrid(range(wf), range(hf)))).reshape(2, -1) + 0.5
    return g / (wf / 2) - 1


def match(img_feat, mesh_feat, mask, pred_v, hf, wf, tau_img=10., tau_mesh=10.):
    bsz, h, w = mask.shape
    mask_down = F.interpolate(mask[:, None], (hf, wf), mode='nearest').reshape(bsz, -1) * 1.0
    pointcorr = mesh_feat.bmm(img_feat).permute(0, 2, 1)
    pointcorr = pointcorr * (mask_down[:, :, None] > 0) - 1e5 * (mask_down[:, :, None] == 0)
    pointcorr_mesh = torch.softmax(tau_mesh * pointcorr, dim=1)
    pointcorr_img = torch.softmax(tau_img * pointcorr, dim=2)
    grid = meshgrid(hf, wf)[None].repeat(bsz, 1, 1).to(pointcorr.dtype)
    imatch = grid.bmm(pointcorr_mesh)
    match3d = (pointcorr_img[:, :, :, None] * pred_v.detach()[:, None, :, :]).sum(2)
    match_up = F.interpolate(match3d.reshape(bsz, hf, wf, 3).permute(0, 3, 1, 2), (h, w), mode='nearest')
    return pointcorr, match_up, imatch, match3d


def rotation_cycle(src_img_feat, tgt_img_feat, src_mask, tgt_mask, cycle_match_gt, hf, wf, tau_mesh=10.):
    """src/tgt_img_feat (B,C,hf*wf) unit-norm, masks (B,1,H,W) (tgt already rotated), gt (B,2,hf/2*wf/2)."""
    bsz, C = src_img_feat.shape[:2]
    grid = meshgrid(hf, wf).reshape(2, hf, wf)[None].repeat(bsz, 1, 1, 1)
    grid = F.interpolate(grid, (hf // 2, wf // 2), mode='bilinear')
    src_mask_down = F.interpolate(src_mask, (hf // 2, wf // 2), mode='nearest').reshape(bsz, -1) * 1.0
    tgt_mask_down = F.interpolate(tgt_mask, (hf // 2, wf // 2), mode='nearest').reshape(bsz, -1) * 1.0
    mask_down = src_mask_down[:, :, None] * tgt_mask_down[:, None, :]
    tgt = F.interpolate(tgt_img_feat.reshape(bsz, C, hf, wf), (hf // 2, wf // 2), mode='nearest').reshape(bsz, C, -1)
    src = F.interpolate(src_img_feat.reshape(bsz, C, hf, wf), (hf // 2, wf // 2), mode='nearest').reshape(bsz, C, -1)
    pointcorr = src.permute(0, 2, 1).bmm(tgt)
    pointcorr = pointcorr * (mask_down > 0) - 1e5 * (mask_down == 0)
    pointcorr_tgt = torch.softmax(tau_mesh * pointcorr, dim=1)
    cycle_match = grid.reshape(bsz, 2, -1).bmm(pointcorr_tgt)
    cycle_loss = ((cycle_match - cycle_match_gt).norm(2, 1) * tgt_mask_down).mean()
    return cycle_loss, cycle_match, tgt_mask_down


def divide_by_frame(x, batch_size, repeat):
    src = x.reshape(batch_size, repeat, *x.shape[1:])
    tgt = torch.cat([src[:, 1:], src[:, :1]], dim=1)
    return src.reshape(-1, *src.shape[2:]), tgt.reshape(-1, *tgt.shape[2:])


def divide_by_instance(x, batch_size, repeat):
    src = x.reshape(batch_size, repeat, *x.shape[1:])
    tgt = torch.cat([src[1:], src[:1]], dim=0)
    return src.reshape(-1, *src.shape[2:]), tgt.reshape(-1, *tgt.shape[2:])


def divide_by_both(x, batch_size, repeat):
    sf, tf = divide_by_frame(x, batch_size, repeat)
    si, ti = divide_by_instance(x, batch_size, repeat)
    return torch.cat([sf, si], dim=0), torch.cat([tf, ti], dim=0)


DIVIDE = {'frame': divide_by_frame, 'instance': divide_by_instance, 'both': divide_by_both}


def pretrain_match(src_feat, tgt_feat, src_mask, tgt_mask, grid, feat_size, k):
    """src/tgt_feat (b,384,fs,fs) DINO keys; masks (b,H,W); grid (b,2,fs,fs)."""
    bsz = src_feat.shape[0]
    src_mask, tgt_mask = src_mask[:, None], tgt_mask[:, None]
    src_feat = src_feat.reshape(*src_feat.shape[:2], -1)
    tgt_feat = tgt_feat.reshape(*tgt_feat.shape[:2], -1)
    src_mask_down = F.interpolate(src_mask, (feat_size, feat_size), mode='nearest').reshape(bsz, -1) * 1.0
    tgt_mask_down = F.interpolate(tgt_mask, (feat_size, feat_size), mode='nearest').reshape(bsz, -1) * 1.0
    mask_down = src_mask_down[:, :, None] * tgt_mask_down[:, None, :]
    pointcorr = src_feat.permute(0, 2, 1).bmm(tgt_feat)
    pointcorr = pointcorr * (mask_down > 0) - 1e5 * (mask_down == 0)
    max_bw = pointcorr.max(1).indices
    max_fw = pointcorr.max(2).indices
    max_cy = torch.gather(max_fw, -1, max_bw)
    grid = grid.reshape(bsz, 2, -1)
    match = torch.gather(grid, -1, max_bw[:, None].repeat(1, 2, 1))
    cycle = torch.gather(grid, -1, max_cy[:, None].repeat(1, 2, 1))
    distance = (cycle - grid).norm(2, 1)
    distance = distance * (tgt_mask_down > 0) + 1e5 * (tgt_mask_down == 0)
    indices = select_topk(-distance, k)
    match = torch.gather(match, -1, indices[:, None].repeat(1, 2, 1))
    grid_k = torch.gather(grid, -1, indices[:, None].repeat(1, 2, 1))
    match_mask = torch.gather(tgt_mask_down, -1, indices)
    indices_match = torch.gather(max_bw, -1, indices)
    return match, grid_k, indices_match, indices, match_mask


def pretrain_cycle_loss(pts_src, indices_tgt, mask_k, depth_weight, pointcorr, batch_size, repeat, hf, wf,
                        divide='both', tau_img=10., tau_mesh=10.):
    """Everything of compute_cycle_loss after self.match(): pointcorr (B,hf*wf,N) full resolution."""
    num_verts = pointcorr.shape[-1]
    fn = DIVIDE[divide]
    dw_src, dw_tgt = fn(depth_weight, batch_size, repeat)
    pc_src, pc_tgt = fn(pointcorr, batch_size, repeat)
    bsz = pc_src.shape[0]
    h2, w2 = hf // 2, wf // 2

    def down(pc):
        return F.interpolate(pc.permute(0, 2, 1).reshape(bsz, num_verts, hf, wf), (h2, w2), mode='bilinear') \
            .reshape(bsz, num_verts, h2 * w2).permute(0, 2, 1)
    pc_src, pc_tgt = down(pc_src), down(pc_tgt)
    pointcorr_img = torch.softmax(tau_img * pc_tgt, dim=2)
    pointcorr_mesh = torch.softmax(tau_mesh * pc_src, dim=1)
    pointcorr_img = pointcorr_img * (dw_tgt[:, None] >= 0.5)
    pointcorr_mesh = pointcorr_mesh * (dw_src[:, None] >= 0.5)
    corr = pointcorr_mesh.bmm(pointcorr_img.permute(0, 2, 1))
    corr = corr / (corr.sum(1, keepdims=True) + 1e-5)
    grid = meshgrid(hf, wf).reshape(2, hf, wf)[None].repeat(bsz, 1, 1, 1)
    grid = F.interpolate(grid, (h2, w2), mode='bilinear').reshape(bsz, 2, -1).to(corr.dtype)
    match = grid.bmm(corr)
    match = torch.gather(match, -1, indices_tgt[:, None].repeat(1, 2, 1))
    cycle_loss = ((match - pts_src).norm(2, 1) * mask_k).mean()
    return cycle_loss, match
