"""CPU: this package's Renderer.render_all (reference launch structure) against golden vectors produced by running the
REFERENCE's own model/module/renderer.py with the CUDA operator replaced by the same deterministic stand-in
(tests/golden/make_renderer_golden.py): pins the host-side chain around the operator -- camera transform, fp64 pinhole
projection, y flip, depth texture, look_at, face gathers, renderer settings / backgrounds, call order, output slicing,
grid_sample and the visibility weight."""
import ast
import os
from types import SimpleNamespace

import numpy as np
import torch

from self_corr_pose_b200.model.module.renderer import Renderer
from self_corr_pose_b200.soft_renderer import modules as sr_modules

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'renderer_golden.npz'))
T = lambda k: torch.from_numpy(np.asarray(G[k]))


def test_render_all_matches_reference_renderer(monkeypatch):
    calls = []

    def fake_rasterize(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
                       eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4,
                       aggr_func_rgb='softmax', aggr_func_alpha='prod', texture_type='surface'):
        B = face_vertices.shape[0]
        calls.append([image_size, list(background_color), near, far, bool(fill_back), eps, sigma_val, dist_func, dist_eps,
                      gamma_val, aggr_func_rgb, aggr_func_alpha, texture_type])
        s = face_vertices.reshape(B, -1).sum(1) * 0.37 + textures.reshape(B, -1).sum(1) * 0.11 + sigma_val * 1e3 + \
            gamma_val * 1e2 + (0.5 if aggr_func_rgb == 'hard' else 0.) + sum(background_color) * 0.25
        ys, xs = torch.meshgrid(torch.arange(image_size, dtype=torch.float32),
                                torch.arange(image_size, dtype=torch.float32), indexing='ij')
        c = torch.arange(4, dtype=torch.float32)[None, :, None, None]
        img = torch.sin(0.1 * (c + 1) * xs[None, None] + 0.07 * ys[None, None] + s[:, None, None, None]) * 0.5 + 0.5
        return img * torch.tensor([1., 1., 6., 1.])[None, :, None, None]

    monkeypatch.setattr(sr_modules.srf, 'soft_rasterize', fake_rasterize)
    opts = SimpleNamespace(img_size=32, use_depth=True)
    mesh = SimpleNamespace(mean_v=T('v'), faces=T('f'), texture_type='vertex')
    outs = Renderer(opts, mesh, reference_launches=True).render_all(
        T('pred_v').clone().requires_grad_(True), T('faces'), T('tex'), T('foc'), T('pp'), T('rot'), T('trans'), None)
    ref_calls = ast.literal_eval(str(G['calls']))
    # same four operator calls; the reference issues mask, soft texture, depth, NOCS -- this package depth, mask, soft
    # texture, NOCS (the order is irrelevant to the results)
    key = lambda c: repr(c)
    assert sorted(calls, key=key) == sorted(ref_calls, key=key)
    names = ('mask_render', 'tex_render', 'depth_render', 'match_gt', 'imatch_gt', 'tex_mask', 'depth_mask', 'match_mask',
             'depth_weight')
    for n, got in zip(names, outs):
        want = T('o_' + n)
        assert got.shape == want.shape, n
        err = float((got.detach() - want).abs().max())
        # the stand-in's phase is a sum over all face coordinates (x 0.37): fp32 summation order differences of the
        # inputs stay below 1e-4 in the image
        assert err <= 2e-4 * max(1.0, float(want.abs().max())), (n, err)
