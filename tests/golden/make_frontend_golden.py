"""Generates tests/golden/frontend_golden.npz by running the REFERENCE's own SoftRas front-end
(third-party/softras/soft_renderer: Mesh, Lighting, Transform/LookAt, SoftRasterizer, SoftRenderer.render_mesh and the
functional helpers) on the CPU in the build container.  The compiled extensions (soft_renderer.cuda.*) and skimage are
stubbed out: the reference package then imports unchanged, and functional.soft_rasterize -- the operator boundary -- is
replaced by a recorder, so the golden vectors are exactly what the reference hands to forward_soft_rasterize
(face_vertices, face_textures and the scalar arguments) for a given mesh / pose / renderer configuration.
Run:  python tests/golden/make_frontend_golden.py      (needs /root/reference; the .npz is committed)
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/third-party/softras'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
for name in ['soft_renderer.cuda', 'soft_renderer.cuda.soft_rasterize', 'soft_renderer.cuda.load_textures',
             'soft_renderer.cuda.create_texture_image', 'soft_renderer.cuda.voxelization', 'skimage', 'skimage.io']:
    sys.modules[name] = types.ModuleType(name)
sys.modules['skimage.io'].imread = sys.modules['skimage.io'].imsave = None
torch.Tensor.cuda = lambda self, *a, **k: self            # mesh.py / look_at.py call .cuda() on constants

import soft_renderer as sr                                  # noqa: E402  (the reference package)
import soft_renderer.rasterizer as ref_rasterizer           # noqa: E402

captured = {}


def recorder(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
             eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb='softmax',
             aggr_func_alpha='prod', texture_type='surface'):
    captured['fv'], captured['ft'] = face_vertices.detach().clone(), textures.detach().clone()
    captured['args'] = [image_size, list(background_color), near, far, bool(fill_back), eps, sigma_val, dist_func, dist_eps,
                        gamma_val, aggr_func_rgb, aggr_func_alpha, texture_type]
    return torch.zeros(face_vertices.shape[0], 4, image_size, image_size)


ref_rasterizer.srf.soft_rasterize = recorder

from self_corr_pose_b200 import synthetic  # noqa: E402

g = torch.Generator().manual_seed(5)
v, f = synthetic.icosphere(1)
B = 2
verts = torch.from_numpy(v)[None].repeat(B, 1, 1) * 0.4 + 0.02 * torch.randn(B, v.shape[0], 3, generator=g)
verts[:, :, 2] += 0.3
faces = torch.from_numpy(f)[None].repeat(B, 1, 1)
vtex = torch.rand(B, v.shape[0], 3, generator=g)
stex = torch.rand(B, f.shape[0], 4, 3, generator=g)
out = dict(verts=verts, faces=faces, vtex=vtex, stex=stex)

CASES = {
    # the model's renderers (model/module/renderer.py:13-26): look_at, orthographic, ambient 1, directional 0
    'model_vertex': dict(ctor=dict(image_size=64, sigma_val=1e-4, gamma_val=1e-4, camera_mode='look_at', perspective=False,
                                   aggr_func_rgb='softmax', light_mode='vertex', light_intensity_ambient=1.,
                                   light_intensity_directionals=0.), tex='vtex', ttype='vertex'),
    'model_mask': dict(ctor=dict(image_size=64, sigma_val=1e-4, gamma_val=1e-4, camera_mode='look_at', perspective=False,
                                 aggr_func_rgb='hard', light_mode='vertex', light_intensity_ambient=1.,
                                 light_intensity_directionals=0.), tex=None, ttype='surface'),
    # library defaults that exercise the rest of the front-end: perspective camera, directional light + normals
    'lit_vertex': dict(ctor=dict(image_size=32, camera_mode='look_at', perspective=True, light_mode='vertex',
                                 light_intensity_ambient=0.4, light_intensity_directionals=0.6,
                                 light_directions=[0.3, 0.8, -0.5]), tex='vtex', ttype='vertex'),
    'lit_surface': dict(ctor=dict(image_size=32, camera_mode='look_at', perspective=True, light_mode='surface',
                                  light_intensity_ambient=0.5, light_intensity_directionals=0.5), tex='stex',
                        ttype='surface'),
}
for name, c in CASES.items():
    r = sr.SoftRenderer(**c['ctor'])
    mesh = sr.Mesh(verts.clone(), faces.clone()) if c['tex'] is None else \
        sr.Mesh(verts.clone(), faces.clone(), out[c['tex']].clone(), texture_type=c['ttype'])
    r.render_mesh(mesh)
    out[name + '_fv'], out[name + '_ft'] = captured['fv'], captured['ft']
    out[name + '_args'] = np.array(repr(captured['args']))
    print(name, tuple(captured['fv'].shape), tuple(captured['ft'].shape), captured['args'])

np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'frontend_golden.npz'),
                    **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
