"""Drop-in `soft_renderer` front-end (the subset reached from MeshNet.forward) on the sm_100a kernels."""
from . import functional
from .mesh import Mesh
from .modules import (AmbientLighting, DirectionalLighting, Lighting, LookAt, Transform, SoftRasterizer,
                      SoftRenderer)

__version__ = '1.0.0-b200'
