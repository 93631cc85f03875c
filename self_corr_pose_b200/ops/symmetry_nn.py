"""Nearest rotated surface sample of every mesh vertex (csrc/scp_sym.cu) as an autograd function: the knn_points(K=1)
step of the reference's one-way chamfer symmetry loss (model/module/mesh.py:53-62, model/util/chamfer.py:152-156) fused
with the sample reconstruction and the rotation.  No CPU path."""
import torch

from .. import _lib


class _SymmetryNN(torch.autograd.Function):

    @staticmethod
    def forward(ctx, pred_v, faces_i32, face_idx, w, rots):
        if not pred_v.is_cuda:
            raise TypeError('symmetry_nn supports only CUDA tensors (no CPU path)')
        B, N, _ = pred_v.shape
        k = rots.shape[0]
        S = face_idx.shape[1]
        pred_v = pred_v.contiguous().float()
        face_idx = face_idx.contiguous()
        w = w.contiguous().float()
        rots = rots.contiguous().float()
        faces_i32 = faces_i32.contiguous()
        _lib.expect_numel('symmetry_nn', face_idx=(face_idx, B * k * S), w=(w, B * k * S * 3), rots=(rots, k * 9))
        if face_idx.dtype != torch.int64 or faces_i32.dtype != torch.int32:
            raise TypeError('symmetry_nn: face_idx must be int64 and faces int32')
        dev = pred_v.device
        dist = torch.empty(B * k, N, device=dev)
        idx = torch.empty(B * k, N, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_symmetry_nn_forward(_lib.ptr(pred_v), _lib.ptr(faces_i32), _lib.ptr(face_idx), _lib.ptr(w),
                                                    _lib.ptr(rots), B, k, N, S, _lib.ptr(dist), _lib.ptr(idx),
                                                    _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_symmetry_nn_forward')
        ctx.save_for_backward(pred_v, faces_i32, face_idx, w, rots, idx)
        ctx.mark_non_differentiable(idx)
        return dist, idx

    @staticmethod
    def backward(ctx, g_dist, _g_idx):
        pred_v, faces_i32, face_idx, w, rots, idx = ctx.saved_tensors
        B, N, _ = pred_v.shape
        k, S = rots.shape[0], face_idx.shape[1]
        dev = pred_v.device
        g = torch.empty_like(pred_v)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_symmetry_nn_backward(_lib.ptr(pred_v), _lib.ptr(faces_i32), _lib.ptr(face_idx), _lib.ptr(w),
                                                     _lib.ptr(rots), _lib.ptr(idx), _lib.ptr(g_dist.contiguous().float()),
                                                     B, k, N, S, _lib.ptr(g), _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_symmetry_nn_backward')
        return g, None, None, None, None


def symmetry_nn(pred_v, faces_i32, face_idx, w, rots):
    """pred_v (B,N,3), faces (nf,3) int32, face_idx (B*k,S) int64, w (B*k,S,3), rots (k,3,3) ->
    (dist (B*k,N) squared distance to the nearest rotated sample, nn_idx (B*k,N) int32)."""
    return _SymmetryNN.apply(pred_v, faces_i32, face_idx, w, rots)
