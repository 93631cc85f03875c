from .geometry import (face_vertices, vertex_normals, look_at, orthogonal, perspective,
                       ambient_lighting, directional_lighting)
from .soft_rasterize import soft_rasterize, SoftRasterizeFunction, soft_rasterize_dual, SoftRasterizeDualFunction
