"""CPU: the torch restatements of the reference's loss / camera functions (self_corr_pose_b200/model/util/loss_utils.py --
the parity reference of the fused kernels) against golden vectors produced by running the REFERENCE's own
model/util/loss_utils.py (tests/golden/make_loss_golden.py), values and gradients."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from self_corr_pose_b200.model.util import loss_utils as L

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'loss_golden.npz'))
T = lambda k: torch.from_numpy(np.asarray(G[k]))


def close(a, b, tol=1e-6):
    a, b = a.double(), b.double()
    assert float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max())), float((a - b).abs().max())


def test_image_losses_match_reference_golden():
    rd, rt, ml = (T(k).clone().requires_grad_(True) for k in ('r_depth', 'r_tex', 'match_lr'))
    img, mask, depth, r_nocs, w = T('img'), T('mask'), T('depth'), T('r_nocs'), T('w')
    B, _, H, _ = rd.shape
    hf = int(round(ml.shape[1] ** 0.5))
    match = F.interpolate(ml.reshape(B, hf, hf, 3).permute(0, 3, 1, 2), (H, H), mode='nearest')
    losses = torch.stack([L.compute_mask_loss(img, mask, rd[:, 3]), L.compute_texture_loss(img, mask, rt[:, :3], rt[:, 3]),
                          L.compute_depth_loss(depth, rd[:, 2], rd[:, 3], mask)[0],
                          L.compute_match_loss(match, r_nocs[:, :3], r_nocs[:, 3], mask)], 1)
    (losses * w).sum().backward()
    close(losses.detach(), T('losses'))
    close(rd.grad, T('g_r_depth'))
    close(rt.grad, T('g_r_tex'))
    close(ml.grad, T('g_match_lr'))
    close(L.compute_imatch_loss(T('imatch'), T('imatch_gt'), T('dw')), T('l_imatch'))


def test_camera_projection_matches_reference_golden():
    leaves = [T(k).clone().requires_grad_(True) for k in ('c_verts', 'c_rot', 'c_trans')]
    sv = L.project_to_screen(leaves[0], T('c_foc'), T('c_pp'), leaves[1], leaves[2])
    (sv * T('c_w')).sum().backward()
    close(sv.detach(), T('c_screen'))
    for leaf, k in zip(leaves, ('c_g_verts', 'c_g_rot', 'c_g_trans')):
        close(leaf.grad, T(k), 1e-5)


def test_laplacian_and_pairing_match_reference_golden():
    lap = L.LaplacianLoss(T('lap_v'), T('lap_f'), average=True)
    x = T('lap_x').clone().requires_grad_(True)
    out = lap(x)
    out.backward()
    close(out.detach(), T('lap_loss'))
    close(x.grad, T('lap_g'))
    t = torch.arange(24.).reshape(12, 2)
    for name in ('divide_by_frame', 'divide_by_instance', 'divide_by_both'):
        s, tt = getattr(L, name)(t, 3, 4)
        assert torch.equal(s, T(name + '_src')) and torch.equal(tt, T(name + '_tgt'))
