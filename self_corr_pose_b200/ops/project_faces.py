"""Fused screen-space geometry (csrc/scp_geom.cu): camera transform + fp64 pinhole projection + y flip + look_at
offset + per-face gathers, forward and backward.

Replaces the torch op chains of the reference's model/util/loss_utils.py:38-61 (pinhole_cam / render),
third-party/softras/soft_renderer/transform.py:29-49 (LookAt of the model's fixed camera) and
functional/face_vertices.py:4-22, which the reference runs once per render.
"""
import math

import torch
from torch.autograd import Function

from .. import _lib

LOOK_AT_Z = 1. / math.tan(math.radians(30.)) + 1.    # soft_renderer/transform.py:33 (viewing_angle 30)


class FaceTopology:
    """faces[nf,3] as int32 on the device + vertex -> face-corner adjacency (CSR), built once per mesh."""

    def __init__(self, faces, num_verts):
        f = faces.detach().reshape(-1, 3).to(torch.int64)
        self.nf, self.N = f.shape[0], int(num_verts)
        flat = f.reshape(-1)
        if flat.numel() and (int(flat.min()) < 0 or int(flat.max()) >= self.N):
            raise ValueError('faces index vertices outside [0, %d)' % self.N)
        order = torch.sort(flat, stable=True).indices                    # corners grouped by vertex, ascending
        counts = torch.bincount(flat, minlength=self.N)
        off = torch.zeros(self.N + 1, dtype=torch.int64, device=f.device)
        off[1:] = torch.cumsum(counts, 0)
        self.faces = f.to(torch.int32).contiguous()
        self.csr_off = off.to(torch.int32).contiguous()
        self.csr_idx = order.to(torch.int32).contiguous()


def check_shapes(pred_v, rotation, translation, foc, pp, topo):
    """(B, N) after checking every buffer against the element count the kernels index."""
    B, N, _ = pred_v.shape
    _lib.expect_numel('project_faces', pred_v=(pred_v, B * N * 3), rotation=(rotation, B * 9),
                      translation=(translation, B * 3), foc=(foc, B * 2), pp=(pp, B * 2))
    if topo is not None and topo.N != N:
        raise ValueError('project_faces: topology built for %d vertices, got %d' % (topo.N, N))
    return B, N


class ProjectFacesFunction(Function):
    """(pred_v[B,N,3], rotation[B,3,3], translation[B,1,3]; foc[B,2] f64, pp[B,2] f64, topo | None) ->
    (screen_v[B,N,3], face_vertices[B,nf,3,3] | None, face_textures[B,nf,3,3] | None)"""

    @staticmethod
    def forward(ctx, pred_v, rotation, translation, foc, pp, topo, z_offset):
        if not pred_v.is_cuda:
            raise TypeError('project_faces supports only CUDA tensors (no CPU path)')
        B, N = check_shapes(pred_v, rotation, translation, foc, pp, topo)
        dev = pred_v.device
        v = pred_v.detach().float().contiguous()
        R = rotation.detach().float().contiguous()
        t = translation.detach().float().reshape(B, 3).contiguous()
        foc = foc.detach().double().contiguous()
        pp = pp.detach().double().contiguous()
        sv = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
        fv = ft = None
        nf = 0
        if topo is not None:
            nf = topo.nf
            fv = torch.empty(B, nf, 3, 3, dtype=torch.float32, device=dev)
            ft = torch.empty(B, nf, 3, 3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_project_faces_forward(
                _lib.ptr(v), _lib.ptr(R), _lib.ptr(t), _lib.ptr(foc), _lib.ptr(pp),
                _lib.ptr(topo.faces) if topo is not None else None, B, N, nf, float(z_offset), _lib.ptr(sv),
                _lib.ptr(fv), _lib.ptr(ft), _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_project_faces_forward')
        ctx.save_for_backward(v, R, t, foc, pp)
        ctx.topo = topo
        ctx.t_shape = translation.shape
        ctx.set_materialize_grads(False)
        return sv, fv, ft

    @staticmethod
    def backward(ctx, g_sv, g_fv, g_ft):
        v, R, t, foc, pp = ctx.saved_tensors
        topo = ctx.topo
        B, N, _ = v.shape
        dev = v.device
        prep = lambda g: None if g is None else g.float().contiguous()
        g_sv, g_fv, g_ft = prep(g_sv), prep(g_fv), prep(g_ft)
        g_v = torch.empty_like(v) if ctx.needs_input_grad[0] else None
        g_R = torch.empty_like(R)
        g_t = torch.empty_like(t)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_project_faces_backward(
                _lib.ptr(v), _lib.ptr(R), _lib.ptr(t), _lib.ptr(foc), _lib.ptr(pp),
                _lib.ptr(topo.csr_off) if topo is not None else None,
                _lib.ptr(topo.csr_idx) if topo is not None else None, B, N, topo.nf if topo is not None else 0,
                _lib.ptr(g_sv), _lib.ptr(g_fv), _lib.ptr(g_ft), _lib.ptr(g_v), _lib.ptr(g_R), _lib.ptr(g_t),
                _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_project_faces_backward')
        return g_v, g_R, g_t.reshape(ctx.t_shape), None, None, None, None


def project_faces(pred_v, rotation, translation, foc, pp, topo=None, z_offset=LOOK_AT_Z):
    return ProjectFacesFunction.apply(pred_v, rotation, translation, foc, pp, topo, z_offset)
