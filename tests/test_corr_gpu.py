"""GPU parity: fused correspondence kernels (through the C ABI / CorrMatchFunction) vs the CPU oracle
(oracle/corr.py, pinned to the reference modules).  Tolerance 1e-3 relative, norm-wise and element-wise."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
import torchvision
from torchvision.transforms import InterpolationMode
from types import SimpleNamespace

from oracle import corr as ocorr

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'corr_golden.npz'))
T = lambda k: torch.from_numpy(np.asarray(G[k]))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def frac(a, b, rtol=1e-3, atol=1e-5):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() <= atol + rtol * b.abs()).double().mean())


def make_inputs(B, hf, wf, N, C=64, H=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    H = H or 4 * hf
    img_feat = F.normalize(torch.randn(B, C, hf * wf, generator=g), 2, 1)
    mesh_feat = F.normalize(torch.relu(torch.randn(B, N, C, generator=g)), 2, -1)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, H), indexing='ij')
    mask = torch.stack([(((xx - 0.1 * b) ** 2 + yy ** 2) < 0.6 ** 2).float() for b in range(B)])
    pred_v = torch.randn(B, N, 3, generator=g)
    return img_feat, mesh_feat, mask, pred_v


def pooled(pc, hf, wf):
    B, P, N = pc.shape
    return F.interpolate(pc.permute(0, 2, 1).reshape(B, N, hf, wf), (hf // 2, wf // 2), mode='bilinear') \
        .reshape(B, N, -1).permute(0, 2, 1)


def oracle_fwd_bwd(img_feat, mesh_feat, mask, pred_v, hf, wf, w_match, w_imatch, w_pool, w_A, dtype=torch.float32):
    img_feat = img_feat.detach().clone().to(dtype).requires_grad_(True)
    mesh_feat = mesh_feat.detach().clone().to(dtype).requires_grad_(True)
    pc, match_up, imatch, match3d = ocorr.match(img_feat, mesh_feat, mask.to(dtype), pred_v.to(dtype), hf, wf)
    pool = pooled(pc, hf, wf)
    # source-side factor of the pre-training cycle loss (pretrained_corr.py:125,131-136): grid . softmax over pixels
    grid2 = F.interpolate(ocorr.meshgrid(hf, wf).reshape(1, 2, hf, wf), (hf // 2, wf // 2), mode='bilinear').reshape(2, -1)
    A = torch.matmul(grid2.to(dtype)[None], torch.softmax(10.0 * pool, dim=1))
    loss = (match3d * w_match.to(dtype)).sum() + (imatch * w_imatch.to(dtype)).sum() + (pool * w_pool.to(dtype)).sum() \
        + (A * w_A.to(dtype)).sum()
    loss.backward()
    return pc.detach(), pool.detach(), match3d.detach(), imatch.detach(), A.detach(), img_feat.grad, mesh_feat.grad


@pytest.mark.parametrize('B,hf,wf,N,bwd', [(2, 16, 16, 70, 'fused'), (2, 32, 32, 1280, 'fused'), (2, 64, 64, 995, 'fused'),
                                           (1, 64, 64, 64, 'fused'), (2, 32, 32, 1280, 'split'), (2, 64, 64, 995, 'split')])
def test_corr_match_forward_backward(B, hf, wf, N, bwd, monkeypatch):
    """bwd = 'fused': the row-block kernel also reduces g_mesh_feat (global float reductions); 'split': separate
    vertex-block kernel (SCP_CORR_BWD=split)."""
    from self_corr_pose_b200.ops.corr_match import corr_match
    monkeypatch.setenv('SCP_CORR_BWD', bwd)
    from self_corr_pose_b200.model.module.correspondence import make_meshgrid
    img_feat, mesh_feat, mask, pred_v = make_inputs(B, hf, wf, N)
    g = torch.Generator().manual_seed(1)
    w_match = torch.randn(B, hf * wf, 3, generator=g)
    w_imatch = torch.randn(B, 2, N, generator=g)
    w_pool = torch.randn(B, hf * wf // 4, N, generator=g) * 0.01
    w_A = torch.randn(B, 2, N, generator=g)
    o = oracle_fwd_bwd(img_feat, mesh_feat, mask, pred_v, hf, wf, w_match, w_imatch, w_pool, w_A)
    o64 = oracle_fwd_bwd(img_feat, mesh_feat, mask, pred_v, hf, wf, w_match, w_imatch, w_pool, w_A, torch.float64)

    mask_down = F.interpolate(mask[:, None], (hf, wf), mode='nearest').reshape(B, -1)
    grid = make_meshgrid(hf, wf, 'cuda')
    torch.testing.assert_close(grid.cpu(), ocorr.meshgrid(hf, wf), rtol=0, atol=0)
    a = img_feat.cuda().requires_grad_(True)
    m = mesh_feat.cuda().requires_grad_(True)
    pc_full, pc_pool, match, imatch, A_pool = corr_match(a, m, mask_down.cuda(), pred_v.cuda(), grid, 10.0, hf, wf,
                                                 want_full=True, want_pool=True)
    loss = (match * w_match.cuda()).sum() + (imatch * w_imatch.cuda()).sum() + (pc_pool * w_pool.cuda()).sum() \
        + (A_pool * w_A.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    got = (pc_full, pc_pool, match, imatch, A_pool, a.grad, m.grad)
    names = ('pointcorr', 'pointcorr_pool', 'match', 'imatch', 'A_pool', 'g_img_feat', 'g_mesh_feat')
    report = {}
    for name, x, ref, ref64 in zip(names, got, o, o64):
        report[name] = (rel(x, ref64), rel(ref, ref64), frac(x, ref))
    print('PARITY corr B%d P%d N%d ' % (B, hf * wf, N) +
          ' '.join('%s=%.2e(oracle32 %.1e, frac %.4f)' % (k, *v) for k, v in report.items()))
    for name, (r, r32, fr) in report.items():
        assert r < 1e-3, (name, r)
        assert fr >= 0.999 or name.startswith('g_'), (name, fr)
    # masked rows are exactly -1e5, like the reference
    assert torch.equal(pc_full.cpu()[o[0] == -1e5], o[0][o[0] == -1e5])


@pytest.mark.parametrize('B,hf,wf,N,want_pool,masks', [
    (2, 16, 16, 70, True, 'disc'), (2, 32, 32, 1280, True, 'disc'), (2, 64, 64, 995, True, 'disc'),
    (1, 64, 64, 64, False, 'disc'), (3, 32, 32, 1024, False, 'empty1'), (3, 64, 64, 1280, True, 'empty1'),
    (2, 64, 64, 1280, True, 'full')])
def test_corr_tc_forward(B, hf, wf, N, want_pool, masks, monkeypatch):
    """Training-mode forward (no full-resolution output) = the tcgen05 path of csrc/scp_corr_tc.cu (similarity on
    kind::tf32 split products with the accumulator in tensor memory, soft-max statistics in the GEMM epilogue): against the
    fp64 oracle, against the mma.sync kernel (SCP_CORR_FWD=legacy) on the same call, and the gradients of the (shared)
    backward from its saved denominators.  'empty1': image 1 has no foreground at all (uniform soft-max fallbacks);
    'full': no background."""
    from self_corr_pose_b200.ops.corr_match import corr_match
    from self_corr_pose_b200.model.module.correspondence import make_meshgrid
    monkeypatch.delenv('SCP_CORR_FWD', raising=False)
    img_feat, mesh_feat, mask, pred_v = make_inputs(B, hf, wf, N, seed=7)
    if masks == 'empty1':
        mask[1] = 0.
    elif masks == 'full':
        mask[:] = 1.
    g = torch.Generator().manual_seed(2)
    w_match = torch.randn(B, hf * wf, 3, generator=g)
    w_imatch = torch.randn(B, 2, N, generator=g)
    w_pool = torch.randn(B, hf * wf // 4, N, generator=g) * 0.01
    w_A = torch.randn(B, 2, N, generator=g)
    if not want_pool:
        w_pool, w_A = w_pool * 0, w_A * 0
    o64 = oracle_fwd_bwd(img_feat, mesh_feat, mask, pred_v, hf, wf, w_match, w_imatch, w_pool, w_A, torch.float64)
    mask_down = F.interpolate(mask[:, None], (hf, wf), mode='nearest').reshape(B, -1).cuda()
    grid = make_meshgrid(hf, wf, 'cuda')

    def run():
        a = img_feat.cuda().requires_grad_(True)
        m = mesh_feat.cuda().requires_grad_(True)
        pc_full, pc_pool, match, imatch, A_pool = corr_match(a, m, mask_down, pred_v.cuda(), grid, 10.0, hf, wf,
                                                             want_full=False, want_pool=want_pool)
        assert pc_full is None and (pc_pool is not None) == want_pool
        loss = (match * w_match.cuda()).sum() + (imatch * w_imatch.cuda()).sum()
        if want_pool:
            loss = loss + (pc_pool * w_pool.cuda()).sum() + (A_pool * w_A.cuda()).sum()
        loss.backward()
        torch.cuda.synchronize()
        return dict(pointcorr_pool=pc_pool, match=match, imatch=imatch, A_pool=A_pool, g_img_feat=a.grad, g_mesh_feat=m.grad)

    got = run()
    monkeypatch.setenv('SCP_CORR_FWD', 'legacy')
    old = run()
    ref = dict(pointcorr_pool=o64[1], match=o64[2], imatch=o64[3], A_pool=o64[4], g_img_feat=o64[5], g_mesh_feat=o64[6])
    line = []
    for k, x in got.items():
        if x is None:
            continue
        r, r_old, r_pair = rel(x, ref[k]), rel(old[k], ref[k]), rel(x, old[k])
        line.append('%s=%.2e(mma.sync %.1e, pair %.1e)' % (k, r, r_old, r_pair))
        if k.startswith('g_'):
            assert r < 1e-3, (k, r)
        else:
            assert r < 1e-5 and frac(x, ref[k].float()) >= 0.999, (k, r)
    print('PARITY corr tcgen05 B%d P%d N%d %s ' % (B, hf * wf, N, masks) + ' '.join(line))
    if want_pool:   # rows of all-background 2x2 blocks are exactly -1e5, like the reference
        assert torch.equal(got['pointcorr_pool'].cpu()[o64[1] == -1e5], o64[1][o64[1] == -1e5].float())


def test_module_match_golden():
    """Correspondence.match module on the golden inputs of the reference run."""
    from self_corr_pose_b200.model.module.correspondence import Correspondence
    opts = SimpleNamespace(tau_img=10., tau_mesh=10., topk_img=100, topk_mesh=100, corr_h=16, corr_w=16,
                           train=True, n_corr_feat=64, img_size=64)
    corr = Correspondence(opts)
    img_feat, mesh_feat, mask, pred_v = (T(k).cuda() for k in ('m_img_feat', 'm_mesh_feat', 'm_mask', 'm_pred_v'))
    pc, match, imatch, conf = corr.match(img_feat, mesh_feat, mask, pred_v)
    assert conf is None
    assert rel(pc, T('m_pointcorr')) < 1e-5 and rel(match, T('m_match')) < 1e-3 and rel(imatch, T('m_imatch')) < 1e-3
    assert frac(match, T('m_match')) >= 0.999 and frac(imatch, T('m_imatch')) >= 0.999


def test_rotation_cycle_vs_oracle():
    """compute_rotation_cycle_loss on a 32x32 map (the kernel needs >= 128 pixels per map; the 16x16 golden
    case pins the oracle on the CPU) against oracle.rotation_cycle, forward and gradients."""
    from self_corr_pose_b200.model.module.correspondence import Correspondence
    hf = wf = 32
    B, C, H = 2, 64, 128
    opts = SimpleNamespace(tau_img=10., tau_mesh=10., topk_img=100, topk_mesh=100, corr_h=hf, corr_w=wf,
                           train=True, n_corr_feat=C, img_size=H)
    corr = Correspondence(opts)
    img_feat, _, mask, _ = make_inputs(B, hf, wf, 8, H=H, seed=4)
    g = torch.Generator().manual_seed(5)
    src_img = torch.rand(B, 3, H, H, generator=g)
    tgt_raw = torch.randn(B, C, hf, wf, generator=g)

    class FakeEncoder:
        def __init__(self, t):
            self.t = t

        def encode_img(self, img, pass_idx=0):
            return None, self.t

    # oracle (CPU)
    torch.manual_seed(11)
    angle = torch.empty(1).uniform_(0., 360.).item()
    rot = torchvision.transforms.functional.rotate
    tgt_mask = rot(mask[:, None], angle, interpolation=InterpolationMode.NEAREST)
    grid = ocorr.meshgrid(hf, wf).reshape(2, hf, wf)[None].repeat(B, 1, 1, 1)
    gt = rot(F.interpolate(grid, (hf // 2, wf // 2), mode='bilinear'), angle,
             interpolation=InterpolationMode.NEAREST).reshape(B, 2, -1)
    a_o = img_feat.clone().requires_grad_(True)
    t_o = tgt_raw.clone().requires_grad_(True)
    loss_o, cm_o, tmd_o = ocorr.rotation_cycle(a_o, F.normalize(t_o.reshape(B, C, -1), 2, 1), mask[:, None], tgt_mask,
                                               gt, hf, wf)
    loss_o.backward()

    a = img_feat.cuda().requires_grad_(True)
    t = tgt_raw.cuda().requires_grad_(True)
    torch.manual_seed(11)
    loss, cm, gt_d, tmd = corr.compute_rotation_cycle_loss(src_img.cuda(), mask.cuda(), a, FakeEncoder(t))
    loss.backward()
    assert torch.equal(gt_d.cpu(), gt) and torch.equal(tmd.cpu(), tmd_o)
    print('PARITY rotcycle loss %.6f vs %.6f cm=%.2e g_src=%.2e g_tgt=%.2e' %
          (float(loss), float(loss_o), rel(cm, cm_o), rel(a.grad, a_o.grad), rel(t.grad, t_o.grad)))
    assert abs(float(loss) - float(loss_o)) < 1e-3 * abs(float(loss_o))
    assert rel(cm, cm_o) < 1e-3 and rel(a.grad, a_o.grad) < 1e-3 and rel(t.grad, t_o.grad) < 1e-3
