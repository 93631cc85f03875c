"""CPU: pins the torch restatements in oracle/corr.py and oracle/vit.py against golden vectors produced by
the reference's own modules (tests/golden/make_corr_golden.py, run where /root/reference is mounted)."""
import os

import numpy as np
import pytest
import torch
import torchvision
from torchvision.transforms import InterpolationMode

from oracle import corr as ocorr
from oracle import vit as ovit
from self_corr_pose_b200.model.module.network.vit_weights import synthetic_state_dict

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'corr_golden.npz'))
T = lambda k: torch.from_numpy(np.asarray(G[k]))
HF = WF = 16


def test_match_golden():
    pc, match, imatch, _ = ocorr.match(T('m_img_feat'), T('m_mesh_feat'), T('m_mask'), T('m_pred_v'), HF, WF)
    torch.testing.assert_close(pc, T('m_pointcorr'), rtol=0, atol=0)
    torch.testing.assert_close(match, T('m_match'), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(imatch, T('m_imatch'), rtol=1e-6, atol=1e-7)


def test_rotation_cycle_golden():
    angle = float(G['r_angle'])
    mask = T('m_mask')[:, None]
    rot = torchvision.transforms.functional.rotate
    tgt_mask = rot(mask, angle, interpolation=InterpolationMode.NEAREST)
    grid = ocorr.meshgrid(HF, WF).reshape(2, HF, WF)[None].repeat(2, 1, 1, 1)
    grid = torch.nn.functional.interpolate(grid, (HF // 2, WF // 2), mode='bilinear')
    gt = rot(grid, angle, interpolation=InterpolationMode.NEAREST).reshape(2, 2, -1)
    torch.testing.assert_close(gt, T('r_cycle_match_gt'), rtol=0, atol=0)
    tgt_feat = torch.nn.functional.normalize(T('r_tgt_feat_raw').reshape(2, 64, -1), 2, 1)
    loss, cm, tmd = ocorr.rotation_cycle(T('m_img_feat'), tgt_feat, mask, tgt_mask, gt, HF, WF)
    torch.testing.assert_close(tmd, T('r_tgt_mask_down'), rtol=0, atol=0)
    torch.testing.assert_close(cm, T('r_cycle_match'), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(loss, T('r_loss'), rtol=1e-6, atol=1e-7)


def test_vit_tiny_golden():
    sd = {k[3:]: T(k) for k in G.files if k.startswith('tw_')}
    k1, _ = ovit.keys_at_layer(sd, T('t_x'), 1, heads=2)
    _, tok2 = ovit.keys_at_layer(sd, T('t_x'), 2, heads=2)
    torch.testing.assert_close(k1, T('t_k1'), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(tok2, T('t_tok2'), rtol=1e-5, atol=1e-6)


def test_dino_features_and_pretrain_cycle_golden():
    sd = synthetic_state_dict(0)
    img = T('p_img')
    feats = ovit.dino_features(sd, img)
    torch.testing.assert_close(feats, T('p_dino_feat'), rtol=1e-4, atol=1e-5)

    bs, rep, k = 2, 2, 20
    mask, dw, pc = T('p_mask'), T('p_depth_weight'), T('p_pointcorr')
    img_src, img_tgt = ocorr.divide_by_both(img, bs, rep)
    m_src, m_tgt = ocorr.divide_by_both(mask, bs, rep)
    f_src, f_tgt = ocorr.divide_by_both(T('p_dino_feat'), bs, rep)
    bsz = f_src.shape[0]
    grid = ocorr.meshgrid(HF, WF).reshape(2, HF, WF)[None].repeat(bsz, 1, 1, 1)
    grid = torch.nn.functional.interpolate(grid, (HF // 2, WF // 2), mode='bilinear')
    pts_src, pts_tgt, idx_src, idx_tgt, mask_k = ocorr.pretrain_match(f_src, f_tgt, m_src, m_tgt, grid, 8, k)
    torch.testing.assert_close(pts_src, T('p_pts_src'), rtol=0, atol=0)
    torch.testing.assert_close(pts_tgt, T('p_pts_tgt'), rtol=0, atol=0)
    torch.testing.assert_close(mask_k, T('p_mask_k'), rtol=0, atol=0)
    loss, match = ocorr.pretrain_cycle_loss(pts_src, idx_tgt, mask_k, dw, pc, bs, rep, HF, WF)
    torch.testing.assert_close(match, T('p_match'), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss, T('p_loss'), rtol=1e-5, atol=1e-7)


def test_product_pretrain_cycle_loss_matches_reference_golden():
    """Host logic of the product's PretrainedCorrespondence.compute_cycle_loss (algebraic elimination of the
    1024x1024 corr matrix, features computed once per unique image) against the reference run, on the CPU,
    with the golden DINO features injected (the ViT itself has no CPU path)."""
    from types import SimpleNamespace
    from self_corr_pose_b200.model.module.pretrained_corr import PretrainedCorrespondence
    opts = SimpleNamespace(tau_img=10., tau_mesh=10., corr_h=HF, corr_w=WF, img_size=64, pretrain_k=20,
                           divide_fn='both', batch_size=2, repeat=2)
    net = PretrainedCorrespondence(opts, mesh=None, device='cpu')
    pc = T('p_pointcorr').clone().requires_grad_(True)
    out = net.compute_cycle_loss(T('p_img'), T('p_mask'), T('p_depth_weight'), pc, feat=T('p_dino_feat'))
    loss, pts_src, pts_tgt, match, mask_k = out[:5]
    torch.testing.assert_close(pts_src, T('p_pts_src'), rtol=0, atol=0)
    torch.testing.assert_close(pts_tgt, T('p_pts_tgt'), rtol=0, atol=0)
    torch.testing.assert_close(mask_k, T('p_mask_k'), rtol=0, atol=0)
    # The reference's fp32 evaluation of corr / (corr.sum + 1e-5) is itself ~6e-4 away from the exact (fp64)
    # value of the same formula; the product must be at least as close to the exact value as the reference is.
    bs, rep = 2, 2
    m_src, m_tgt = ocorr.divide_by_both(T('p_mask'), bs, rep)
    f_src, f_tgt = ocorr.divide_by_both(T('p_dino_feat'), bs, rep)
    grid = ocorr.meshgrid(HF, WF).reshape(2, HF, WF)[None].repeat(f_src.shape[0], 1, 1, 1)
    grid = torch.nn.functional.interpolate(grid, (HF // 2, WF // 2), mode='bilinear')
    ps, pt, i_s, i_t, mk = ocorr.pretrain_match(f_src, f_tgt, m_src, m_tgt, grid, 8, 20)
    loss64, match64 = ocorr.pretrain_cycle_loss(ps.double(), i_t, mk.double(), T('p_depth_weight').double(),
                                                T('p_pointcorr').double(), bs, rep, HF, WF)
    ref_err = (T('p_match').double() - match64).abs().max()
    my_err = (match.double() - match64).abs().max()
    assert my_err <= max(2 * ref_err, 1e-5), (my_err, ref_err)
    assert abs(float(loss) - float(loss64)) <= max(2 * abs(float(G['p_loss']) - float(loss64)), 1e-6)
    # gradient w.r.t. pointcorr equals the reference formulation's (oracle autograd, fp64)
    loss.backward()
    pc2 = T('p_pointcorr').double().requires_grad_(True)
    l2, _ = ocorr.pretrain_cycle_loss(ps.double(), i_t, mk.double(), T('p_depth_weight').double(), pc2, bs, rep, HF, WF)
    l2.backward()
    pc3 = T('p_pointcorr').clone().requires_grad_(True)     # the reference formulation in fp32 (what it actually runs)
    l3, _ = ocorr.pretrain_cycle_loss(ps, i_t, mk, T('p_depth_weight'), pc3, bs, rep, HF, WF)
    l3.backward()
    gerr = (pc.grad.double() - pc2.grad).norm() / pc2.grad.norm()
    gerr_ref = (pc3.grad.double() - pc2.grad).norm() / pc2.grad.norm()
    assert gerr <= max(2 * gerr_ref, 1e-3), (gerr, gerr_ref)

    # pooled input path gives the same loss
    pooled = torch.nn.functional.avg_pool2d(T('p_pointcorr').permute(0, 2, 1).reshape(4, -1, HF, WF), 2) \
        .reshape(4, -1, (HF // 2) * (WF // 2)).permute(0, 2, 1)
    out2 = net.compute_cycle_loss(T('p_img'), T('p_mask'), T('p_depth_weight'), pooled, pooled=True,
                                  feat=T('p_dino_feat'))
    assert abs(float(out2[0]) - float(loss64)) <= max(2 * abs(float(G['p_loss']) - float(loss64)), 1e-6)
