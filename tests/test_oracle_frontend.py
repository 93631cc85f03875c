"""CPU: this package's SoftRas front-end (soft_renderer: Mesh, Lighting, Transform/LookAt, SoftRasterizer,
SoftRenderer.render_mesh, functional helpers) against golden vectors produced by running the REFERENCE's own
third-party/softras/soft_renderer front-end up to the operator boundary (tests/golden/make_frontend_golden.py): the
face_vertices / face_textures tensors and the scalar arguments handed to soft_rasterize must be the same."""
import ast
import os

import numpy as np
import pytest
import torch

from self_corr_pose_b200 import soft_renderer as sr
from self_corr_pose_b200.soft_renderer import modules as sr_modules

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'frontend_golden.npz'))
T = lambda k: torch.from_numpy(np.asarray(G[k]))

CASES = {
    'model_vertex': dict(ctor=dict(image_size=64, sigma_val=1e-4, gamma_val=1e-4, camera_mode='look_at', perspective=False,
                                   aggr_func_rgb='softmax', light_mode='vertex', light_intensity_ambient=1.,
                                   light_intensity_directionals=0.), tex='vtex', ttype='vertex'),
    'model_mask': dict(ctor=dict(image_size=64, sigma_val=1e-4, gamma_val=1e-4, camera_mode='look_at', perspective=False,
                                 aggr_func_rgb='hard', light_mode='vertex', light_intensity_ambient=1.,
                                 light_intensity_directionals=0.), tex=None, ttype='surface'),
    'lit_vertex': dict(ctor=dict(image_size=32, camera_mode='look_at', perspective=True, light_mode='vertex',
                                 light_intensity_ambient=0.4, light_intensity_directionals=0.6,
                                 light_directions=[0.3, 0.8, -0.5]), tex='vtex', ttype='vertex'),
    'lit_surface': dict(ctor=dict(image_size=32, camera_mode='look_at', perspective=True, light_mode='surface',
                                  light_intensity_ambient=0.5, light_intensity_directionals=0.5), tex='stex',
                        ttype='surface'),
}


@pytest.mark.parametrize('name', list(CASES))
def test_frontend_hands_the_operator_what_the_reference_does(name, monkeypatch):
    c = CASES[name]
    captured = {}

    def recorder(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
                 eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb='softmax',
                 aggr_func_alpha='prod', texture_type='surface'):
        captured['fv'], captured['ft'] = face_vertices.detach().clone(), textures.detach().clone()
        captured['args'] = [image_size, list(background_color), near, far, bool(fill_back), eps, sigma_val, dist_func,
                            dist_eps, gamma_val, aggr_func_rgb, aggr_func_alpha, texture_type]
        return torch.zeros(face_vertices.shape[0], 4, image_size, image_size)

    monkeypatch.setattr(sr_modules.srf, 'soft_rasterize', recorder)
    r = sr.SoftRenderer(**c['ctor'])
    verts, faces = T('verts'), T('faces')
    mesh = sr.Mesh(verts.clone(), faces.clone()) if c['tex'] is None else \
        sr.Mesh(verts.clone(), faces.clone(), T(c['tex']).clone(), texture_type=c['ttype'])
    r.render_mesh(mesh)
    assert captured['args'] == ast.literal_eval(str(G[name + '_args']))
    for key, got in (('_fv', captured['fv']), ('_ft', captured['ft'])):
        want = T(name + key)
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max())), (name, key)
