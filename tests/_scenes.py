"""Shared synthetic SoftRas scenes for the tests (SURVEY.md section 8d, config 0 and friends)."""
import numpy as np
import torch

from self_corr_pose_b200 import synthetic
from self_corr_pose_b200.model.util.loss_utils import project_to_screen
from self_corr_pose_b200.soft_renderer import functional as srf

# the four renderer configurations of model/module/renderer.py:13-27 (+ background colours)
RENDER_CONFIGS = {
    'mask':    dict(sigma_val=1e-4, gamma_val=1e-4, aggr_func_rgb='hard', background_color=(0, 0, 0)),
    'softtex': dict(sigma_val=1e-3, gamma_val=1e-2, aggr_func_rgb='softmax', background_color=(1, 1, 1)),
    'depth':   dict(sigma_val=1e-4, gamma_val=1e-4, aggr_func_rgb='softmax', background_color=(1, 1, 1)),
    'hardtex': dict(sigma_val=1e-4, gamma_val=1e-3, aggr_func_rgb='hard', background_color=(0, 0, 0)),
}
LOOK_AT_Z = 1. / np.tan(np.radians(30.)) + 1.


def mesh(name):
    if name == 'uv1280':
        return synthetic.uv_sphere()
    if name == 'ico642':
        return synthetic.icosphere(3)
    return synthetic.load_prior(name)


def screen_faces(verts, faces, rot, trans, foc=3.7):
    """numpy mesh + pose -> (face_vertices [B,nf,3,3], screen verts [B,V,3]) exactly as the renderer
    front-end produces them (project, y-flip, look_at offset, orthographic)."""
    B = rot.shape[0]
    v = torch.from_numpy(verts)[None].repeat(B, 1, 1)
    f = torch.from_numpy(faces)[None].repeat(B, 1, 1)
    foc_t = torch.full((B, 2), foc, dtype=torch.float64)
    pp_t = torch.zeros(B, 2, dtype=torch.float64)
    sv = project_to_screen(v, foc_t, pp_t, rot, trans)
    sv = srf.orthogonal(srf.look_at(sv, [0, 0, -LOOK_AT_Z]), 1.0)
    return srf.face_vertices(sv, f), sv, f


def config0(mesh_name, B=1, seed=0):
    """R = Rx(0.6) Ry(0.7), t = (0,0,5) for image 0; further images get seeded random poses."""
    verts, faces = mesh(mesh_name)
    g = torch.Generator().manual_seed(seed)
    rot, trans = synthetic.random_poses(B, g)
    rot[0] = torch.from_numpy(synthetic.rot_xyz(0.6, 0.7, 0.0))
    trans[0] = torch.tensor([[0., 0., 5.]])
    return screen_faces(verts, faces, rot, trans)


def vertex_colors(sv, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(sv.shape, generator=g)
