"""CPU: pins oracle/softras_oracle.c against the reference's own kernel source compiled for the
host (oracle/_ref/libsoftras_ref_cpu.so), on the config-0 scenes for the four
renderer configurations of the model, forward and backward, plus the other operator modes."""
import numpy as np
import pytest
import torch

from oracle import softras as osr
from tests import _scenes
from self_corr_pose_b200.soft_renderer import functional as srf

pytestmark = pytest.mark.skipif(not osr.have_ref(), reason='reference CPU build unavailable')


def _textures(kind, sv, f, fv):
    if kind == 'mask':      # sr.Mesh(verts, faces): surface texture of ones, R = 1
        return np.ones((fv.shape[0], fv.shape[1], 1, 3), np.float32), 'surface'
    if kind == 'depth':     # texture = screen-space vertices
        return srf.face_vertices(sv, f).numpy(), 'vertex'
    return srf.face_vertices(_scenes.vertex_colors(sv), f).numpy(), 'vertex'


@pytest.mark.parametrize('mesh_name', ['laptop', 'uv1280'])
@pytest.mark.parametrize('kind', ['mask', 'softtex', 'depth', 'hardtex'])
def test_oracle_matches_reference_build(mesh_name, kind):
    fv, sv, f = _scenes.config0(mesh_name, B=2)
    cfg = dict(_scenes.RENDER_CONFIGS[kind])
    tex, ttype = _textures(kind, sv, f, fv)
    kw = dict(image_size=64, texture_type=ttype, **cfg)
    col_o, info_o, aggr_o = osr.forward(fv.numpy(), tex, **kw)
    col_r, info_r, aggr_r = osr.ref_forward(fv.numpy(), tex, **kw)
    # same compiler, same arithmetic order, -ffp-contract=off on both: expect bit-level agreement
    np.testing.assert_allclose(info_o, info_r, rtol=0, atol=0)
    np.testing.assert_allclose(col_o, col_r, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(aggr_o, aggr_r, rtol=1e-6, atol=1e-7)
    assert col_o[:, 3].max() > 0.99 and col_o[:, 3].min() < 1e-3   # the object is actually in view

    g = np.random.RandomState(3).randn(*col_o.shape).astype(np.float32)
    gf_o, gt_o = osr.backward(fv.numpy(), tex, col_o, info_o, aggr_o, g, nthreads=1, **kw)
    gf_r, gt_r = osr.ref_backward(fv.numpy(), tex, col_r, info_r, aggr_r, g, **kw)
    scale = np.abs(gf_r).max()
    np.testing.assert_allclose(gf_o, gf_r, rtol=1e-4, atol=1e-5 * scale)
    np.testing.assert_allclose(gt_o, gt_r, rtol=1e-4, atol=1e-5 * max(np.abs(gt_r).max(), 1e-6))
    assert scale > 0


@pytest.mark.parametrize('dist_func,alpha', [('barycentric', 'sum'), ('hard', 'hard'), ('euclidean', 'sum')])
def test_other_modes_match_reference_build(dist_func, alpha):
    fv, sv, f = _scenes.config0('ico642', B=1)
    tex = srf.face_vertices(_scenes.vertex_colors(sv), f).numpy()
    kw = dict(image_size=32, texture_type='vertex', sigma_val=1e-3, gamma_val=1e-2, aggr_func_rgb='softmax',
              dist_func=dist_func, aggr_func_alpha=alpha, background_color=(0.2, 0.4, 0.6))
    col_o, info_o, aggr_o = osr.forward(fv.numpy(), tex, **kw)
    col_r, info_r, aggr_r = osr.ref_forward(fv.numpy(), tex, **kw)
    np.testing.assert_allclose(col_o, col_r, rtol=1e-6, atol=1e-7)
    g = np.random.RandomState(5).randn(*col_o.shape).astype(np.float32)
    gf_o, gt_o = osr.backward(fv.numpy(), tex, col_o, info_o, aggr_o, g, nthreads=1, **kw)
    gf_r, gt_r = osr.ref_backward(fv.numpy(), tex, col_r, info_r, aggr_r, g, **kw)
    np.testing.assert_allclose(gf_o, gf_r, rtol=1e-4, atol=1e-5 * max(np.abs(gf_r).max(), 1e-6))
    np.testing.assert_allclose(gt_o, gt_r, rtol=1e-4, atol=1e-5 * max(np.abs(gt_r).max(), 1e-6))


def test_surface_texture_table():
    """R = 2 surface textures (texel lookup path).  Hard RGB only: with soft RGB, pixels outside a
    triangle get clipped weights of exactly (1,0,0), for which the reference reads one texel past the
    face's table (soft_rasterize_cuda_kernel.cu:182-187, i.e. the NEXT face's texel); the oracle and
    the sm_100a kernel clamp the index instead (documented deviation)."""
    fv, sv, f = _scenes.config0('ico642', B=1)
    nf = fv.shape[1]
    tex = np.random.RandomState(0).rand(1, nf, 4, 3).astype(np.float32)
    kw = dict(image_size=32, texture_type='surface', sigma_val=1e-4, gamma_val=1e-3, aggr_func_rgb='hard')
    col_o, info_o, aggr_o = osr.forward(fv.numpy(), tex, **kw)
    col_r, info_r, aggr_r = osr.ref_forward(fv.numpy(), tex, **kw)
    np.testing.assert_allclose(col_o, col_r, rtol=1e-6, atol=1e-7)
    g = np.random.RandomState(5).randn(*col_o.shape).astype(np.float32)
    gf_o, gt_o = osr.backward(fv.numpy(), tex, col_o, info_o, aggr_o, g, nthreads=1, **kw)
    gf_r, gt_r = osr.ref_backward(fv.numpy(), tex, col_r, info_r, aggr_r, g, **kw)
    np.testing.assert_allclose(gt_o, gt_r, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gf_o, gf_r, rtol=1e-4, atol=1e-5 * np.abs(gf_r).max())
