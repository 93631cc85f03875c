"""Batched triangle-mesh container with lazily gathered per-face attributes.

API of third-party/softras/soft_renderer/mesh.py:9-134 (constructor arguments, `vertices`,
`faces`, `textures`, `face_vertices`, `face_textures`, `surface_normals`, `vertex_normals`,
`fill_back_`, `reset_`).  OBJ IO and voxelisation are outside the hot path and not provided.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import functional as srf


def _to_tensor(x, dtype):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x).to(dtype)
        if torch.cuda.is_available():
            x = x.cuda()
    return x


class Mesh(object):
    def __init__(self, vertices, faces, textures=None, texture_res=1, texture_type='surface'):
        vertices = _to_tensor(vertices, torch.float32)
        faces = _to_tensor(faces, torch.int32)
        if vertices.ndimension() == 2:
            vertices = vertices[None]
        if faces.ndimension() == 2:
            faces = faces[None]
        self._vertices, self._faces = vertices, faces
        self.device = vertices.device
        self.texture_type = texture_type
        self.batch_size = vertices.shape[0]
        self.num_vertices = vertices.shape[1]
        self.num_faces = faces.shape[1]
        self._cache = {}
        self._fill_back = False

        if textures is None:
            if texture_type == 'surface':
                textures = torch.ones(self.batch_size, self.num_faces, texture_res ** 2, 3,
                                      dtype=torch.float32, device=self.device)
                self.texture_res = texture_res
            elif texture_type == 'vertex':
                textures = torch.ones(self.batch_size, self.num_vertices, 3, dtype=torch.float32,
                                      device=self.device)
                self.texture_res = 1
            else:
                raise ValueError('texture type not applicable')
        else:
            textures = _to_tensor(textures, torch.float32)
            if textures.ndimension() == 3 and texture_type == 'surface':
                textures = textures[None]
            if textures.ndimension() == 2 and texture_type == 'vertex':
                textures = textures[None]
            self.texture_res = int(np.sqrt(textures.shape[2]))
        self._textures = textures
        self._origin = (self._vertices, self._faces, self._textures)

    # -- mutable geometry: any change invalidates the gathered caches
    @property
    def vertices(self):
        return self._vertices

    @vertices.setter
    def vertices(self, v):
        self._vertices = v
        self.num_vertices = v.shape[1]
        self._cache.clear()

    @property
    def faces(self):
        return self._faces

    @faces.setter
    def faces(self, f):
        self._faces = f
        self.num_faces = f.shape[1]
        self._cache.clear()

    @property
    def textures(self):
        return self._textures

    @textures.setter
    def textures(self, t):
        self._textures = t

    def _cached(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    @property
    def face_vertices(self):
        return self._cached('fv', lambda: srf.face_vertices(self.vertices, self.faces))

    @property
    def surface_normals(self):
        def compute():
            fv = self.face_vertices
            return F.normalize(torch.cross(fv[:, :, 2] - fv[:, :, 1], fv[:, :, 0] - fv[:, :, 1], dim=-1),
                               p=2, dim=2, eps=1e-6)
        return self._cached('sn', compute)

    @property
    def vertex_normals(self):
        return self._cached('vn', lambda: srf.vertex_normals(self.vertices, self.faces))

    @property
    def face_textures(self):
        if self.texture_type == 'surface':
            return self.textures
        if self.texture_type == 'vertex':
            return srf.face_vertices(self.textures, self.faces)
        raise ValueError('texture type not applicable')

    def fill_back_(self):
        if not self._fill_back:
            self.faces = torch.cat((self.faces, self.faces[:, :, [2, 1, 0]]), dim=1)
            self.textures = torch.cat((self.textures, self.textures), dim=1)
            self._fill_back = True

    def reset_(self):
        self.vertices, self.faces, self.textures = self._origin
        self._fill_back = False
