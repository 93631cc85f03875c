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
