"""GPU parity: the sm_100a SoftRas kernels (through the C ABI / SoftRasterizeFunction) against the
CPU oracle on identical inputs.  Tolerance 1e-3 relative (BASELINE.json north_star), stated per check;
outputs that go through hard decisions (z-buffer winner, inside/outside flips under 1-ulp changes)
are required to agree on >= 99.9 % of the elements."""
import numpy as np
import pytest
import torch

from oracle import softras as osr
from tests import _scenes
from self_corr_pose_b200.soft_renderer import functional as srf

pytestmark = pytest.mark.gpu

RTOL = 1e-3


def frac_close(a, b, rtol=RTOL, atol=1e-5, env=None):
    tol = atol + rtol * np.abs(b)
    if env is not None:
        tol = tol + env
    return float(np.mean(np.abs(a - b) <= tol))


def rel_norm(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-12))


def oracle_pair(fv, tex, g_seed, kw):
    """Oracle forward+backward twice: strict fp32 (pinned to the reference build) and with FMA
    contraction.  The reference's Gram-matrix distance formula is ill-conditioned near silhouette
    edges (fragment exponent = d^2/sigma with cancellation in d): two valid fp32 evaluations of the
    SAME algorithm differ by up to 1e-2 there, so |strict - fma| is used as a per-element envelope."""
    res = []
    g = None
    for fma in (False, True):
        col, info, aggr = osr.forward(fv.numpy(), tex.numpy(), fma=fma, **kw)
        if g is None:
            g = np.random.RandomState(g_seed).randn(*col.shape).astype(np.float32)
        gf, gt = osr.backward(fv.numpy(), tex.numpy(), col, info, aggr, g, fma=fma, **kw)
        res.append((col, gf, gt))
    return g, res[0], res[1]


def check_against_oracle(tag, got, strict, fma, k_env=8.0, min_frac=0.999, rtol_norm=RTOL):
    """got/strict/fma = (colors, grad_faces, grad_textures).

    Primary check: against the FMA-contracted oracle build (same rounding model as nvcc's default, which
    is also how the reference's own kernel is compiled) at the bare 1e-3 tolerance.
    Secondary check: against the strict (-ffp-contract=off, pinned bit-for-bit to the reference source
    built for the CPU) oracle with the |strict - fma| conditioning envelope."""
    col, gf, gt = got
    col_o, gf_o, gt_o = strict
    col_f, gf_f, gt_f = fma
    gt_o, gt_f = gt_o.reshape(gt.shape), gt_f.reshape(gt.shape)
    env = k_env * np.abs(col_o - col_f)
    has_gt = np.abs(gt_f).max() > 0
    m = dict(
        alpha=frac_close(col[:, 3], col_f[:, 3]), rgb=frac_close(col[:, :3], col_f[:, :3]),
        col_rel=rel_norm(col, col_f), gf_rel=rel_norm(gf, gf_f),
        gt_rel=rel_norm(gt, gt_f) if has_gt else float(np.abs(gt).max()),
        alpha_env=frac_close(col[:, 3], col_o[:, 3], env=env[:, 3]),
        rgb_env=frac_close(col[:, :3], col_o[:, :3], env=env[:, :3]),
        col_rel_strict=rel_norm(col, col_o), oracles_col_rel=rel_norm(col_f, col_o),
        gf_rel_strict=rel_norm(gf, gf_o), oracles_gf_rel=rel_norm(gf_f, gf_o))
    print('PARITY %s %s' % (tag, ' '.join('%s=%.3g' % kv for kv in m.items())))
    assert m['alpha'] >= min_frac and m['rgb'] >= min_frac, m
    assert m['col_rel'] <= rtol_norm and m['gf_rel'] <= rtol_norm and m['gt_rel'] <= rtol_norm, m
    assert m['alpha_env'] >= min_frac and m['rgb_env'] >= min_frac, m
    assert m['col_rel_strict'] <= max(rtol_norm, 3 * m['oracles_col_rel']), m
    assert m['gf_rel_strict'] <= max(rtol_norm, 3 * m['oracles_gf_rel']), m
    return m


def _textures(kind, sv, f, fv):
    if kind == 'mask':
        return torch.ones(fv.shape[0], fv.shape[1], 1, 3), 'surface'
    if kind == 'depth':
        return srf.face_vertices(sv, f), 'vertex'
    return srf.face_vertices(_scenes.vertex_colors(sv), f), 'vertex'


def run_gpu(fv, tex, g, kw):
    fv_d = fv.cuda().requires_grad_(True)
    tex_d = tex.cuda().requires_grad_(True)
    out = srf.soft_rasterize(fv_d, tex_d, **kw)
    out.backward(torch.from_numpy(g).cuda())
    torch.cuda.synchronize()
    return out.detach().cpu().numpy(), fv_d.grad.cpu().numpy(), tex_d.grad.cpu().numpy()


@pytest.mark.parametrize('mesh_name,size,B', [('laptop', 64, 2), ('uv1280', 64, 1), ('laptop', 256, 2),
                                              ('ico642', 100, 3)])
@pytest.mark.parametrize('kind', ['mask', 'softtex', 'depth', 'hardtex'])
def test_forward_backward_vs_oracle(mesh_name, size, B, kind):
    fv, sv, f = _scenes.config0(mesh_name, B=B)
    cfg = dict(_scenes.RENDER_CONFIGS[kind])
    tex, ttype = _textures(kind, sv, f, fv)
    kw = dict(image_size=size, texture_type=ttype, **cfg)
    g, strict, fma = oracle_pair(fv, tex, 3, kw)
    got = run_gpu(fv, tex, g, kw)
    check_against_oracle('%s/%s/%d' % (kind, mesh_name, size), got, strict, fma)


@pytest.mark.parametrize('dist_func,alpha,rgb', [('barycentric', 'sum', 'softmax'), ('hard', 'hard', 'hard'),
                                                ('euclidean', 'sum', 'softmax'), ('euclidean', 'hard', 'softmax')])
def test_generic_modes_vs_oracle(dist_func, alpha, rgb):
    fv, sv, f = _scenes.config0('ico642', B=2)
    tex = srf.face_vertices(_scenes.vertex_colors(sv), f)
    kw = dict(image_size=48, texture_type='vertex', sigma_val=1e-3, gamma_val=1e-2, aggr_func_rgb=rgb,
              dist_func=dist_func, aggr_func_alpha=alpha, background_color=(0.2, 0.4, 0.6))
    g, strict, fma = oracle_pair(fv, tex, 5, kw)
    got = run_gpu(fv, tex, g, kw)
    check_against_oracle('%s/%s/%s' % (dist_func, alpha, rgb), got, strict, fma)


def test_surface_table_and_ragged_size():
    """R = 2 surface textures, image size not a multiple of the 16-pixel tile.  The texel index is
    int(w*R): a hard decision that flips under 1-ulp changes of the clipped weights and then returns an
    unrelated (random) texel, so this (non-hot-path) mode is held to >= 99 % of elements / 5e-2 norm-wise."""
    fv, sv, f = _scenes.config0('ico642', B=1)
    tex = torch.rand(1, fv.shape[1], 4, 3, generator=torch.Generator().manual_seed(0))
    kw = dict(image_size=37, texture_type='surface', sigma_val=1e-4, gamma_val=1e-3, aggr_func_rgb='softmax')
    g, strict, fma = oracle_pair(fv, tex, 5, kw)
    got = run_gpu(fv, tex, g, kw)
    check_against_oracle('surface_R2', got, strict, fma, min_frac=0.99, rtol_norm=5e-2)


def test_empty_view_and_single_face():
    """Mesh entirely outside the image -> background untouched, zero gradients; one-face mesh works."""
    fv, sv, f = _scenes.config0('ico642', B=1)
    far_away = fv.clone()
    far_away[..., 0] += 10.0
    tex = torch.ones(1, fv.shape[1], 1, 3)
    kw = dict(image_size=32, texture_type='surface', sigma_val=1e-4, gamma_val=1e-4, aggr_func_rgb='hard',
              background_color=(0.25, 0.5, 0.75))
    g = np.ones((1, 4, 32, 32), np.float32)
    col, gf, gt = run_gpu(far_away, tex, g, kw)
    assert np.all(col[:, 3] == 0) and np.all(col[:, 0] == 0.25) and np.all(col[:, 2] == 0.75)
    assert np.abs(gf).max() == 0
    one = fv[:, :1].clone()
    col_o, info_o, aggr_o = osr.forward(one.numpy(), np.ones((1, 1, 1, 3), np.float32), **kw)
    col, gf, gt = run_gpu(one, torch.ones(1, 1, 1, 3), g, kw)
    assert frac_close(col, col_o) >= 0.999


def test_render_is_deterministic_forward():
    fv, sv, f = _scenes.config0('laptop', B=2)
    tex = srf.face_vertices(_scenes.vertex_colors(sv), f)
    kw = dict(image_size=128, texture_type='vertex', **_scenes.RENDER_CONFIGS['softtex'])
    a = srf.soft_rasterize(fv.cuda(), tex.cuda(), **kw)
    b = srf.soft_rasterize(fv.cuda(), tex.cuda(), **kw)
    assert torch.equal(a, b)


@pytest.mark.parametrize('mesh,size,B', [('uv1280', 128, 3), ('ico642', 256, 2), ('laptop', 64, 2)])
def test_dual_traversal_equals_two_renders(mesh, size, B):
    """scp_softras_forward_dual (depth render + NOCS map in one traversal) against the two separate launches of the
    same kernels: forward outputs identical (up to FMA contraction), gradient of the soft render equal up to atomic-order noise."""
    fv, sv, f = _scenes.config0(mesh, B=B)
    tex_soft = srf.face_vertices(sv, f).cuda()
    tex_hard = srf.face_vertices(_scenes.vertex_colors(sv) - 0.5, f).cuda()
    cfg_d, cfg_h = _scenes.RENDER_CONFIGS['depth'], _scenes.RENDER_CONFIGS['hardtex']
    a = fv.cuda().requires_grad_(True)
    ts = tex_soft.clone().requires_grad_(True)
    r_soft = srf.soft_rasterize(a, ts, image_size=size, texture_type='vertex', **cfg_d)
    r_hard = srf.soft_rasterize(a.detach(), tex_hard, image_size=size, texture_type='vertex', **cfg_h)
    g = torch.randn(r_soft.shape, generator=torch.Generator().manual_seed(5)).cuda()
    r_soft.backward(g)
    a2 = fv.cuda().requires_grad_(True)
    ts2 = tex_soft.clone().requires_grad_(True)
    d_soft, d_hard = srf.soft_rasterize_dual(a2, ts2, tex_hard, image_size=size, sigma_val=cfg_d['sigma_val'],
                                             gamma_val=cfg_d['gamma_val'], background_soft=cfg_d['background_color'],
                                             background_hard=cfg_h['background_color'])
    d_soft.backward(g)
    torch.cuda.synchronize()
    # same statements in both instantiations; only the compiler's FMA contraction choices may differ
    assert torch.allclose(d_soft, r_soft, rtol=1e-5, atol=1e-6)
    assert float((d_hard != r_hard).float().mean()) < 1e-4 and torch.equal(d_hard[:, 3], d_soft[:, 3])
    assert not d_hard.requires_grad
    for x, y, name in ((a2.grad, a.grad, 'faces'), (ts2.grad, ts.grad, 'textures')):
        rel = float((x - y).norm() / y.norm().clamp_min(1e-30))
        print('PARITY dual-vs-separate grad_%s %s %dpx rel=%.2e' % (name, mesh, size, rel))
        assert rel < 1e-4


@pytest.mark.parametrize('kind', ['softtex', 'depth', 'hardtex'])
def test_backward_face_centric_equals_tile_centric(kind, monkeypatch):
    """The two backward traversals (default: one warp per face, no atomics; SCP_SOFTRAS_BWD=tile: tile-centric with
    global reductions) evaluate the same per-(pixel, face) terms: gradients agree up to summation order."""
    fv, sv, f = _scenes.config0('uv1280', B=3)
    tex = srf.face_vertices(_scenes.vertex_colors(sv), f).cuda()
    kw = dict(image_size=128, texture_type='vertex', **_scenes.RENDER_CONFIGS[kind])
    g = torch.randn(3, 4, 128, 128, generator=torch.Generator().manual_seed(9)).cuda()
    grads = {}
    for mode in ('face', 'tile'):
        monkeypatch.setenv('SCP_SOFTRAS_BWD', mode)
        a = fv.cuda().requires_grad_(True)
        t = tex.clone().requires_grad_(True)
        srf.soft_rasterize(a, t, **kw).backward(g)
        torch.cuda.synchronize()
        grads[mode] = (a.grad.clone(), t.grad.clone())
    # hard RGB: the geometry gradient is the alpha term alone, ill-conditioned in fp32 (exponent d^2 / sigma with
    # cancellation in d; the two oracle builds differ by 2-25 % on it) -- the instantiations differ by FMA contraction
    # choices there, which is why the product keeps the oracle-pinned tile kernel for hard RGB by default
    tol = 5e-2 if kind == 'hardtex' else 1e-4
    for name, x, y in zip(('faces', 'textures'), grads['face'], grads['tile']):
        rel = float((x - y).norm() / y.norm().clamp_min(1e-30))
        print('PARITY face-vs-tile backward %s grad_%s rel=%.2e' % (kind, name, rel))
        assert rel < (tol if name == 'faces' else 1e-4)
    # the face-centric traversal has a single writer per face: bit-reproducible run to run
    monkeypatch.setenv('SCP_SOFTRAS_BWD', 'face')
    a = fv.cuda().requires_grad_(True)
    t = tex.clone().requires_grad_(True)
    srf.soft_rasterize(a, t, **kw).backward(g)
    assert torch.equal(a.grad, grads['face'][0]) and torch.equal(t.grad, grads['face'][1])
