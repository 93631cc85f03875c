"""nn.Module front-end of the soft renderer: lighting, camera transform, rasteriser, renderer.

API of third-party/softras/soft_renderer/{lighting.py:8-69, transform.py:29-112,
rasterizer.py:9-56, renderer.py:47-105}.  Camera modes used by the hot path ('look_at',
orthographic or perspective) are implemented; 'projection' / 'look' raise.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as srf
from .mesh import Mesh


class AmbientLighting(nn.Module):
    def __init__(self, light_intensity=0.5, light_color=(1, 1, 1)):
        super().__init__()
        self.light_intensity, self.light_color = light_intensity, light_color

    def forward(self, light):
        return srf.ambient_lighting(light, self.light_intensity, self.light_color)


class DirectionalLighting(nn.Module):
    def __init__(self, light_intensity=0.5, light_color=(1, 1, 1), light_direction=(0, 1, 0)):
        super().__init__()
        self.light_intensity, self.light_color, self.light_direction = light_intensity, light_color, light_direction

    def forward(self, light, normals):
        return srf.directional_lighting(light, normals, self.light_intensity, self.light_color,
                                        self.light_direction)


class Lighting(nn.Module):
    def __init__(self, light_mode='surface', intensity_ambient=0.5, color_ambient=[1, 1, 1],
                 intensity_directionals=0.5, color_directionals=[1, 1, 1], directions=[0, 1, 0]):
        super().__init__()
        if light_mode not in ('surface', 'vertex'):
            raise ValueError('Lighting mode only support surface and vertex')
        self.light_mode = light_mode
        self.ambient = AmbientLighting(intensity_ambient, color_ambient)
        self.directionals = nn.ModuleList([DirectionalLighting(intensity_directionals, color_directionals,
                                                               directions)])

    def forward(self, mesh):
        surface = self.light_mode == 'surface'
        like = mesh.faces if surface else mesh.vertices
        light = torch.zeros(like.shape, dtype=torch.float32, device=mesh.device)
        light = self.ambient(light)
        for d in self.directionals:
            # a zero-intensity directional light adds exactly 0 (renderer.py:15 of the model uses
            # intensity 0): skip the normal computation, result unchanged
            if d.light_intensity == 0:
                continue
            light = d(light, mesh.surface_normals if surface else mesh.vertex_normals)
        mesh.textures = mesh.textures * (light[:, :, None, :] if surface else light)
        return mesh


class LookAt(nn.Module):
    def __init__(self, perspective=True, viewing_angle=30, viewing_scale=1.0, eye=None):
        super().__init__()
        self.perspective, self.viewing_angle, self.viewing_scale = perspective, viewing_angle, viewing_scale
        self._eye = eye if eye is not None else [0, 0, -(1. / math.tan(math.radians(viewing_angle)) + 1)]

    def forward(self, vertices):
        vertices = srf.look_at(vertices, self._eye)
        if self.perspective:
            return srf.perspective(vertices, angle=self.viewing_angle)
        return srf.orthogonal(vertices, scale=self.viewing_scale)


class Transform(nn.Module):
    def __init__(self, camera_mode='projection', P=None, dist_coeffs=None, orig_size=512, perspective=True,
                 viewing_angle=30, viewing_scale=1.0, eye=None, camera_direction=[0, 0, 1]):
        super().__init__()
        self.camera_mode = camera_mode
        if camera_mode == 'look_at':
            self.transformer = LookAt(perspective, viewing_angle, viewing_scale, eye)
        elif camera_mode in ('projection', 'look'):
            raise NotImplementedError("camera mode '%s' is outside the self-corr-pose hot path" % camera_mode)
        else:
            raise ValueError('Camera mode has to be one of projection, look or look_at')

    def forward(self, mesh):
        mesh.vertices = self.transformer(mesh.vertices)
        return mesh

    def set_eyes(self, eyes):
        self.transformer._eye = eyes


class SoftRasterizer(nn.Module):
    def __init__(self, image_size=256, background_color=[0, 0, 0], near=1, far=100, anti_aliasing=False,
                 fill_back=False, eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4,
                 aggr_func_rgb='softmax', aggr_func_alpha='prod', texture_type='surface'):
        super().__init__()
        if dist_func not in ('hard', 'euclidean', 'barycentric'):
            raise ValueError('Distance function only support hard, euclidean and barycentric')
        if aggr_func_rgb not in ('hard', 'softmax'):
            raise ValueError('Aggregate function(rgb) only support hard and softmax')
        if aggr_func_alpha not in ('hard', 'prod', 'sum'):
            raise ValueError('Aggregate function(a) only support hard, prod and sum')
        if texture_type not in ('surface', 'vertex'):
            raise ValueError('Texture type only support surface and vertex')
        self.image_size, self.background_color = image_size, background_color
        self.near, self.far, self.eps = near, far, eps
        self.anti_aliasing, self.fill_back = anti_aliasing, fill_back
        self.sigma_val, self.gamma_val = sigma_val, gamma_val
        self.dist_func, self.dist_eps = dist_func, dist_eps
        self.aggr_func_rgb, self.aggr_func_alpha = aggr_func_rgb, aggr_func_alpha
        self.texture_type = texture_type

    def forward(self, mesh, mode=None):
        size = self.image_size * (2 if self.anti_aliasing else 1)
        images = srf.soft_rasterize(mesh.face_vertices, mesh.face_textures, size, self.background_color,
                                    self.near, self.far, self.fill_back, self.eps, self.sigma_val,
                                    self.dist_func, self.dist_eps, self.gamma_val, self.aggr_func_rgb,
                                    self.aggr_func_alpha, self.texture_type)
        if self.anti_aliasing:
            images = F.avg_pool2d(images, kernel_size=2, stride=2)
        return images


class SoftRenderer(nn.Module):
    def __init__(self, image_size=256, background_color=[0, 0, 0], near=1, far=100, anti_aliasing=False,
                 fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func='euclidean', dist_eps=1e-4, gamma_val=1e-4,
                 aggr_func_rgb='softmax', aggr_func_alpha='prod', texture_type='surface',
                 camera_mode='projection', P=None, dist_coeffs=None, orig_size=512, perspective=True,
                 viewing_angle=30, viewing_scale=1.0, eye=None, camera_direction=[0, 0, 1],
                 light_mode='surface', light_intensity_ambient=0.5, light_color_ambient=[1, 1, 1],
                 light_intensity_directionals=0.5, light_color_directionals=[1, 1, 1],
                 light_directions=[0, 1, 0]):
        super().__init__()
        self.lighting = Lighting(light_mode, light_intensity_ambient, light_color_ambient,
                                 light_intensity_directionals, light_color_directionals, light_directions)
        self.transform = Transform(camera_mode, P, dist_coeffs, orig_size, perspective, viewing_angle,
                                   viewing_scale, eye, camera_direction)
        self.rasterizer = SoftRasterizer(image_size, background_color, near, far, anti_aliasing, fill_back, eps,
                                         sigma_val, dist_func, dist_eps, gamma_val, aggr_func_rgb,
                                         aggr_func_alpha, texture_type)

    def set_sigma(self, sigma):
        self.rasterizer.sigma_val = sigma

    def set_gamma(self, gamma):
        self.rasterizer.gamma_val = gamma

    def set_texture_mode(self, mode):
        assert mode in ('vertex', 'surface'), 'Mode only support surface and vertex'
        self.lighting.light_mode = mode
        self.rasterizer.texture_type = mode

    def render_mesh(self, mesh, mode=None):
        self.set_texture_mode(mesh.texture_type)
        mesh = self.lighting(mesh)
        mesh = self.transform(mesh)
        return self.rasterizer(mesh, mode)

    def forward(self, vertices, faces, textures=None, mode=None, texture_type='surface'):
        return self.render_mesh(Mesh(vertices, faces, textures=textures, texture_type=texture_type), mode)
