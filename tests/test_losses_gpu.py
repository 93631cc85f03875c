"""GPU parity of the fused image-space losses (csrc/scp_loss.cu through ops/image_losses.py) against the reference's
op-by-op statements (model/util/loss_utils.py compute_mask_loss / compute_texture_loss / compute_depth_loss /
compute_match_loss, restated in self_corr_pose_b200/model/util/loss_utils.py), values and gradients."""
import pytest
import torch
import torch.nn.functional as F

from self_corr_pose_b200.model.util import loss_utils as L

pytestmark = pytest.mark.gpu


def _inputs(B, H, hf, seed, dev='cuda'):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    r_depth = torch.cat([r(B, 2, H, H), 3 * r(B, 1, H, H), (r(B, 1, H, H) - 0.3).clamp(0, 1)], 1)
    r_tex = r(B, 4, H, H)
    r_nocs = torch.cat([r(B, 3, H, H) - 0.5, (r(B, 1, H, H) - 0.4).clamp(0, 1)], 1)
    match_lr = r(B, hf * hf, 3) - 0.5
    img = r(B, 3, H, H)
    mask = (r(B, H, H) > 0.4).float()
    depth = 3 * r(B, H, H) * (r(B, H, H) > 0.2).float()
    w = r(B, 4) + 0.5
    return [t.to(dev) for t in (r_depth, r_tex, r_nocs, match_lr, img, mask, depth, w)]


def _reference(r_depth, r_tex, r_nocs, match_lr, img, mask, depth, hf, use_depth):
    B, _, H, W = r_depth.shape
    match = F.interpolate(match_lr.reshape(B, hf, hf, 3).permute(0, 3, 1, 2), (H, W), mode='nearest')
    lm = L.compute_mask_loss(img, mask, r_depth[:, 3])
    lt = L.compute_texture_loss(img, mask, r_tex[:, :3], r_tex[:, 3])
    ld = L.compute_depth_loss(depth, r_depth[:, 2], r_depth[:, 3], mask)[0] if use_depth else torch.zeros_like(lm)
    lmt = L.compute_match_loss(match, r_nocs[:, :3], r_nocs[:, 3], mask)
    return torch.stack([lm, lt, ld, lmt], 1)


@pytest.mark.parametrize('B,H,hf,use_depth', [(3, 64, 16, True), (2, 128, 32, True), (2, 256, 64, True), (2, 64, 64, False),
                                              (1, 48, 16, True)])
def test_image_losses_match_reference_statements(B, H, hf, use_depth):
    from self_corr_pose_b200.ops.image_losses import image_losses
    r_depth, r_tex, r_nocs, match_lr, img, mask, depth, w = _inputs(B, H, hf, seed=B + H)
    leaves = [t.clone().requires_grad_(True) for t in (r_depth, r_tex, match_lr)]
    ref = _reference(leaves[0].double(), leaves[1].double(), r_nocs.double(), leaves[2].double(), img.double(),
                     mask.double(), depth.double(), hf, use_depth)
    (ref * w.double()).sum().backward()
    g_ref = [t.grad.clone() for t in leaves]
    for t in leaves:
        t.grad = None
    out = torch.stack(image_losses(leaves[0], leaves[1], leaves[2], img, mask, depth, r_nocs, hf, hf, use_depth), 1)
    (out * w).sum().backward()
    torch.cuda.synchronize()
    for k, name in enumerate(('mask', 'texture', 'depth', 'match')):
        err = float((out[:, k].double() - ref[:, k]).abs().max() / ref[:, k].abs().max().clamp_min(1e-30))
        print('PARITY image_losses %s B%d H%d value rel=%.2e' % (name, B, H, err))
        assert err < 1e-5 or (k == 2 and not use_depth)
    for t, gr, name in zip(leaves, g_ref, ('r_depth', 'r_tex', 'match_lr')):
        err = float((t.grad.double() - gr).norm() / gr.norm().clamp_min(1e-30))
        print('PARITY image_losses grad %s B%d H%d rel=%.2e' % (name, B, H, err))
        assert err < 1e-5


def test_hotpath_step_fused_equals_op_by_op():
    """Whole step: fused losses + shared geometry vs the reference's op-by-op statements on the same kernels."""
    from self_corr_pose_b200 import synthetic
    from self_corr_pose_b200.hotpath import HotPath, default_opts
    from self_corr_pose_b200.model.module.renderer import Renderer
    opts = default_opts(img_size=128, corr_h=32, corr_w=32, batch_size=2, repeat=2, pretrain_k=50)
    v, f = synthetic.icosphere(3)
    res = []
    for fused in (True, False):
        hot = HotPath(opts, torch.from_numpy(v), torch.from_numpy(f), device='cuda', fused_losses=fused)
        data, enc = synthetic.make_batch(opts, v, f, 4, device='cuda', seed=3, renderer=Renderer(opts, hot.mesh))
        total, aux = hot.step(data, enc)
        res.append((float(total), {k: float(x) for k, x in aux.items()}, [e.grad.clone() for e in enc]))
    (ta, auxa, ga), (tb, auxb, gb) = res
    for k in auxb:
        assert abs(auxa[k] - auxb[k]) <= 1e-4 * abs(auxb[k]) + 1e-8, (k, auxa[k], auxb[k])
    for name, a, b in zip(('img_feat', 'mesh_feat', 'pred_v', 'rotation', 'translation'), ga, gb):
        rel = float((a - b).norm() / b.norm().clamp_min(1e-30))
        print('PARITY fused-vs-op-by-op grad %s rel=%.2e' % (name, rel))
        assert rel < 2e-3   # SoftRas backward accumulates with atomics in both runs (order noise)


@pytest.mark.parametrize('B,P4,N,k', [(4, 256, 642, 50), (6, 1024, 1280, 200)])
def test_cycle_rows_match_reference_statements(B, P4, N, k):
    """ops/cycle_rows.py (csrc/scp_cycle.cu) against the op-by-op statements of pretrained_corr.py:120-139."""
    from self_corr_pose_b200.ops.cycle_rows import cycle_rows
    g = torch.Generator().manual_seed(B + k)
    NP = 2 * B
    pc = (torch.rand(B, P4, N, generator=g) * 2 - 1)
    pc[:, ::7] -= 25000.          # rows of partially masked pooled blocks
    A = torch.rand(B, 2, N, generator=g) * 2 - 1
    dw = torch.rand(B, N, generator=g)
    src_idx = torch.arange(NP) % B
    tgt_idx = (torch.arange(NP) + 1 + (torch.arange(NP) // B)) % B
    rows = torch.stack([torch.randperm(P4, generator=g)[:k] for _ in range(NP)])
    pts = torch.rand(NP, 2, k, generator=g) * 2 - 1
    mask_k = (torch.rand(NP, k, generator=g) > 0.3).float()
    w = torch.rand(NP, generator=g) + 0.5
    pc, A, dw, src_idx, tgt_idx, rows, pts, mask_k, w = (t.cuda() for t in (pc, A, dw, src_idx, tgt_idx, rows, pts, mask_k, w))
    tau = 10.0

    pc_r, A_r = pc.double().requires_grad_(True), A.double().requires_grad_(True)
    dw_src, dw_tgt = dw.index_select(0, src_idx), dw.index_select(0, tgt_idx)
    A_src = A_r.index_select(0, src_idx) * (dw_src[:, None] >= 0.5)
    s_src = (dw_src >= 0.5).double()
    flat = (tgt_idx[:, None] * P4 + rows).reshape(-1)
    r = pc_r.reshape(-1, N).index_select(0, flat).reshape(NP, k, N)
    Pi = torch.softmax(tau * r, dim=2) * (dw_tgt[:, None] >= 0.5)
    match_r = torch.matmul(A_src, Pi.permute(0, 2, 1)) / (torch.matmul(s_src[:, None], Pi.permute(0, 2, 1)) + 1e-5)
    pair_r = ((match_r - pts.double()).norm(2, 1) * mask_k.double()).sum(1)
    (pair_r * w.double()).sum().backward()

    pc_g, A_g = pc.clone().requires_grad_(True), A.clone().requires_grad_(True)
    pair, match = cycle_rows(pc_g, A_g, dw, src_idx, tgt_idx, rows, pts, mask_k, tau)
    (pair * w).sum().backward()
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    res = dict(pair_loss=rel(pair, pair_r), match=rel(match, match_r), g_pc=rel(pc_g.grad, pc_r.grad), g_A=rel(A_g.grad, A_r.grad))
    print('PARITY cycle_rows B%d P4 %d N%d k%d ' % (B, P4, N, k) + ' '.join('%s=%.2e' % kv for kv in res.items()))
    for name, v in res.items():
        assert v < 1e-4, (name, v)


def test_fused_kernels_match_reference_golden():
    """Fused image losses and fused geometry on the golden inputs of tests/golden/make_loss_golden.py: outputs of the
    REFERENCE's own loss_utils functions (values and autograd gradients)."""
    import os
    import numpy as np
    from self_corr_pose_b200.ops.image_losses import image_losses
    from self_corr_pose_b200.ops.project_faces import project_faces
    from self_corr_pose_b200.model.util.loss_utils import LaplacianLoss
    G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'loss_golden.npz'))
    T = lambda k: torch.from_numpy(np.asarray(G[k])).cuda()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    rd, rt, ml = (T(k).clone().requires_grad_(True) for k in ('r_depth', 'r_tex', 'match_lr'))
    hf = int(round(ml.shape[1] ** 0.5))
    out = torch.stack(image_losses(rd, rt, ml, T('img'), T('mask'), T('depth'), T('r_nocs'), hf, hf, True), 1)
    (out * T('w')).sum().backward()
    res = dict(losses=rel(out, T('losses')), g_r_depth=rel(rd.grad, T('g_r_depth')), g_r_tex=rel(rt.grad, T('g_r_tex')),
               g_match_lr=rel(ml.grad, T('g_match_lr')))
    leaves = [T(k).clone().requires_grad_(True) for k in ('c_verts', 'c_rot', 'c_trans')]
    sv = project_faces(leaves[0], leaves[1], leaves[2], T('c_foc'), T('c_pp'))[0]
    (sv * T('c_w')).sum().backward()
    res.update(screen=rel(sv, T('c_screen')), g_verts=rel(leaves[0].grad, T('c_g_verts')),
               g_rot=rel(leaves[1].grad, T('c_g_rot')), g_trans=rel(leaves[2].grad, T('c_g_trans')))
    lap = LaplacianLoss(T('lap_v').cpu(), T('lap_f').cpu(), average=True).cuda()
    x = T('lap_x').clone().requires_grad_(True)
    ll = lap(x)
    ll.backward()
    res.update(lap=rel(ll, T('lap_loss')), lap_g=rel(x.grad, T('lap_g')))
    print('PARITY fused kernels vs reference golden ' + ' '.join('%s=%.2e' % kv for kv in res.items()))
    for k, v in res.items():
        assert v < 1e-5, (k, v)
