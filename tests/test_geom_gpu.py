"""GPU parity of the fused screen-space geometry (csrc/scp_geom.cu through ops/project_faces.py) against the
reference's op chain (loss_utils.project_to_screen = pinhole_cam + y flip, softras look_at + orthogonal,
face_vertices), values and gradients."""
import pytest
import torch

from self_corr_pose_b200 import synthetic
from self_corr_pose_b200.model.util.loss_utils import project_to_screen
from self_corr_pose_b200.soft_renderer import functional as srf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,mesh', [(3, 'ico'), (2, 'uv')])
def test_project_faces_matches_op_chain(B, mesh):
    from self_corr_pose_b200.ops.project_faces import project_faces, FaceTopology, LOOK_AT_Z
    v, f = synthetic.icosphere(2) if mesh == 'ico' else synthetic.uv_sphere()
    g = torch.Generator().manual_seed(B)
    rot, trans = synthetic.random_poses(B, g)
    foc = (3.7 + 0.3 * torch.rand(B, 2, generator=g)).double().cuda()
    pp = (0.1 * (torch.rand(B, 2, generator=g) - 0.5)).double().cuda()
    N, nf = v.shape[0], f.shape[0]
    pv0 = (torch.from_numpy(v)[None] + 0.01 * torch.randn(B, N, 3, generator=g)).cuda()
    faces = torch.from_numpy(f).cuda()
    fb = faces[None].repeat(B, 1, 1)
    w_sv, w_fv, w_ft = (torch.randn(s, generator=g).cuda() for s in ((B, N, 3), (B, nf, 3, 3), (B, nf, 3, 3)))

    leaves_ref = [t.clone().double().requires_grad_(True) for t in (pv0, rot.cuda(), trans.cuda())]
    sv_r = project_to_screen(leaves_ref[0], foc, pp, leaves_ref[1], leaves_ref[2])
    # look_at from (0,0,-LOOK_AT_Z) towards the origin with up = y is the identity rotation (checked in fp32 below)
    off = torch.tensor([0., 0., LOOK_AT_Z], dtype=torch.float64, device='cuda')
    fv_r = srf.face_vertices(sv_r + off, fb)
    probe = torch.randn(1, 5, 3, generator=g).cuda()
    assert torch.allclose(srf.orthogonal(srf.look_at(probe, [0, 0, -LOOK_AT_Z]), 1.0),
                          probe + off.float(), rtol=0, atol=1e-6)
    ft_r = srf.face_vertices(sv_r, fb)
    ((sv_r * w_sv).sum() + (fv_r * w_fv).sum() + (ft_r * w_ft).sum()).backward()

    leaves = [t.clone().requires_grad_(True) for t in (pv0, rot.cuda(), trans.cuda())]
    sv, fv, ft = project_faces(leaves[0], leaves[1], leaves[2], foc, pp, FaceTopology(faces, N))
    ((sv * w_sv).sum() + (fv * w_fv).sum() + (ft * w_ft).sum()).backward()
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    for name, a, b in (('screen_v', sv, sv_r), ('face_vertices', fv, fv_r), ('face_textures', ft, ft_r)):
        print('PARITY project_faces %s rel=%.2e' % (name, rel(a, b)))
        assert rel(a, b) < 1e-6
    for name, a, b in zip(('pred_v', 'rotation', 'translation'), leaves, leaves_ref):
        print('PARITY project_faces grad %s rel=%.2e' % (name, rel(a.grad, b.grad)))
        assert rel(a.grad, b.grad) < 1e-5
    # projection only (imatch_gt path): no faces, pred_v detached
    sv2 = project_faces(pv0, leaves[1], leaves[2], foc, pp)[0]
    assert torch.equal(sv2, sv.detach())


def test_laplacian_loss_sparse_equals_dense():
    """LaplacianLoss on CUDA (sparse product, scp_spmm3) against the reference's dense-buffer matmul
    (loss_utils.py:63-97), value and gradient."""
    from self_corr_pose_b200.model.util.loss_utils import LaplacianLoss
    v, f = synthetic.uv_sphere()
    mod = LaplacianLoss(torch.from_numpy(v), torch.from_numpy(f), average=True)
    g = torch.Generator().manual_seed(0)
    x = torch.from_numpy(v)[None] + 0.05 * torch.randn(5, v.shape[0], 3, generator=g)
    x_ref = x.clone().double().requires_grad_(True)
    ref = (torch.matmul(mod.laplacian.double(), x_ref).pow(2).sum((1, 2))).sum() / x.shape[0]
    ref.backward()
    mod = mod.cuda()
    x_d = x.cuda().requires_grad_(True)
    out = mod(x_d)
    out.backward()
    rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))
    print('PARITY laplacian value rel=%.2e grad rel=%.2e' % (rel(out, ref), rel(x_d.grad, x_ref.grad)))
    assert rel(out, ref) < 1e-5 and rel(x_d.grad, x_ref.grad) < 1e-5
