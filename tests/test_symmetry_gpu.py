"""GPU parity of the fused symmetry regulariser (csrc/scp_sym.cu: sample reconstruction + rotation + 1-NN) against the
reference's statements op by op (CanonicalMesh.compute_symmetry_loss_reference = model/module/mesh.py:53-62 +
model/util/chamfer.py:152-221 of the reference with a brute-force knn) on the SAME random draws.  Tolerance 1e-5
relative on the loss, 1e-4 norm-wise on the gradient (a nearest-neighbour tie broken differently moves one term)."""
import pytest
import torch

from self_corr_pose_b200.hotpath import default_opts
from self_corr_pose_b200.model.module.mesh import CanonicalMesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,npts,symmetry_idx', [(3, 700, 1), (2, 10000, 1), (2, 1000, 0), (1, 513, 2)])
def test_fused_symmetry_loss_matches_reference_statements(B, npts, symmetry_idx):
    opts = default_opts(symmetry_idx=symmetry_idx)
    mesh = CanonicalMesh(opts).cuda()
    g = torch.Generator().manual_seed(B * 1000 + npts)
    pred_v = (mesh.mean_v.detach().cpu()[None] + 0.03 * torch.randn(B, mesh.num_verts, 3, generator=g)).cuda()
    faces = mesh.faces[None].repeat(B, 1, 1)
    out = {}
    for name, fn in (('fused', mesh.compute_symmetry_loss), ('ref', mesh.compute_symmetry_loss_reference)):
        v = pred_v.clone().requires_grad_(True)
        torch.manual_seed(7)
        torch.cuda.manual_seed(7)
        loss = fn(v, faces, npts)
        loss.backward()
        out[name] = (float(loss), v.grad.clone())
    rel_loss = abs(out['fused'][0] - out['ref'][0]) / abs(out['ref'][0])
    rel_grad = float((out['fused'][1] - out['ref'][1]).norm() / out['ref'][1].norm())
    print('PARITY symmetry B%d S%d k-idx%d loss=%.6e rel_loss=%.2e rel_grad=%.2e' % (B, npts, symmetry_idx, out['ref'][0],
                                                                                  rel_loss, rel_grad))
    assert rel_loss < 1e-5
    assert rel_grad < 1e-4
