"""GPU <-> GPU parity: this package's SoftRas operator against "R-GPU", the reference's own legacy CUDA kernel
(third-party/softras/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu:674-813) recompiled by nvcc for sm_100a
(baseline/build_ref_gpu.py).  nvcc's compilation of the reference source is the arbiter of the fp32 conditioning
of the reference algorithm.  Skipped when baseline/_ref is absent.

Tolerance: 1e-3 relative norm-wise on colours and gradients, >= 99.9 % of the pixels within 1e-3 (north star).
The legacy backward accumulates with atomicAdd in launch order, so its own gradients carry order noise."""
import numpy as np
import pytest
import torch

from tests import _rgpu, _scenes
from self_corr_pose_b200.soft_renderer import functional as srf

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _rgpu.available(), reason='baseline/_ref (R-GPU) not built')]


def _tex(kind, sv, f, fv):
    if kind == 'mask':
        return torch.ones(fv.shape[0], fv.shape[1], 1, 3), 'surface'
    if kind == 'depth':
        return srf.face_vertices(sv, f), 'vertex'
    return srf.face_vertices(_scenes.vertex_colors(sv), f), 'vertex'


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _frac(a, b, rtol=1e-3, atol=1e-5):
    return float(((a - b).abs() <= atol + rtol * b.abs()).float().mean())


# configs[0] (64 px, laptop + the 1280-vertex sphere) and configs[2]-shaped inputs (256 px, uv1280, B = 4)
@pytest.mark.parametrize('mesh_name,size,B', [('laptop', 64, 1), ('uv1280', 64, 1), ('uv1280', 256, 4), ('laptop', 256, 2)])
@pytest.mark.parametrize('kind', ['mask', 'softtex', 'depth', 'hardtex'])
def test_operator_vs_legacy_kernel(mesh_name, size, B, kind):
    fv, sv, f = _scenes.config0(mesh_name, B=B)
    tex, ttype = _tex(kind, sv, f, fv)
    cfg = dict(_scenes.RENDER_CONFIGS[kind])
    kw = dict(image_size=size, texture_type=ttype, **cfg)
    fv_d, tex_d = fv.cuda(), tex.cuda()
    g = torch.randn(B, 4, size, size, generator=torch.Generator().manual_seed(3)).cuda()

    col_r, info_r, aggr_r = _rgpu.forward(fv_d, tex_d, **kw)
    gf_r, gt_r = _rgpu.backward(fv_d, tex_d, col_r, info_r, aggr_r, g, **kw)

    a = fv_d.clone().requires_grad_(True)
    t = tex_d.clone().requires_grad_(True)
    col = srf.soft_rasterize(a, t, **kw)
    col.backward(g)
    torch.cuda.synchronize()
    m = dict(alpha=_frac(col[:, 3], col_r[:, 3]), rgb=_frac(col[:, :3], col_r[:, :3]), col_rel=_rel(col, col_r),
             gf_rel=_rel(a.grad, gf_r),
             gt_rel=_rel(t.grad, gt_r) if float(gt_r.abs().max()) > 0 else float(t.grad.abs().max()))
    print('PARITY-RGPU %s/%s/%d %s' % (kind, mesh_name, size, ' '.join('%s=%.3g' % kv for kv in m.items())))
    assert m['alpha'] >= 0.999 and m['rgb'] >= 0.999, m
    assert m['col_rel'] <= 1e-3 and m['gf_rel'] <= 1e-3 and m['gt_rel'] <= 1e-3, m
