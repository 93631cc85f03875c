"""Camera projection, render wrapper and the loss terms of the training step.

Same names, arguments and arithmetic as the reference's model/util/loss_utils.py
(pinhole_cam :38-47, render :49-61, LaplacianLoss :63-97, compute_mask_loss :236-244,
compute_texture_loss :246-252, compute_depth_loss :273-284, compute_match_loss :317-320,
compute_imatch_loss :322-324, divide_by_frame/instance/both :326-345).  Pure torch (device
agnostic); the rasterisation inside `render` goes to the sm_100a SoftRas kernels.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import soft_renderer as sr


def pinhole_cam(verts, pp, foc):
    """In-place pinhole projection x' = pp + x*f/z (fp64 intrinsics promote, result cast back into verts)."""
    if verts.dim() == 3:
        z = verts[:, :, 2].clone()
        for ax in (1, 0):
            verts[:, :, ax] = pp[:, ax][:, None] + verts[:, :, ax].clone() * foc[:, ax][:, None] / z
    elif verts.dim() == 2:
        z = verts[:, 2].clone()
        for ax in (1, 0):
            verts[:, ax] = pp[ax] + verts[:, ax].clone() * foc[ax] / z
    else:
        raise ValueError("vertices shape must be (bsz, N, 3) or (N, 3).")
    return verts


def project_to_screen(verts, foc, pp, rotation, translation):
    """verts.bmm(R) + t -> pinhole -> y-flip: screen-space vertices fed to SoftRas (render :54-57)."""
    verts = verts.bmm(rotation) + translation
    verts = pinhole_cam(verts, pp, foc)
    verts[:, :, 1] *= -1
    return verts


def render(renderer, verts, faces, tex, foc, pp, rotation, translation, rotation_detach=False,
           translation_detach=False, render_depth=False, render_mask=False, texture_type='vertex'):
    rot = rotation.clone().detach() if rotation_detach else rotation.clone()
    trans = translation.clone().detach() if translation_detach else translation.clone()
    verts = project_to_screen(verts, foc, pp, rot, trans)
    if render_depth:
        tex = verts.clone()
    if render_mask:
        return renderer.render_mesh(sr.Mesh(verts, faces))
    return renderer.render_mesh(sr.Mesh(verts, faces, tex, texture_type=texture_type))


class LaplacianLoss(nn.Module):
    """Uniform-weight graph Laplacian smoothness (dense N x N buffer, rows normalised by the degree)."""

    def __init__(self, vertex, faces, average=False):
        super().__init__()
        self.nv, self.nf, self.average = vertex.size(0), faces.size(0), average
        f = faces.detach().cpu().numpy() if torch.is_tensor(faces) else np.asarray(faces)
        lap = np.zeros([self.nv, self.nv], np.float32)
        for a, b in ((0, 1), (1, 2), (2, 0)):
            lap[f[:, a], f[:, b]] = -1
            lap[f[:, b], f[:, a]] = -1
        deg = -lap.sum(1)
        lap[np.arange(self.nv), np.arange(self.nv)] = deg
        nz = deg != 0
        lap[nz] /= deg[nz, None]
        self.register_buffer('laplacian', torch.from_numpy(lap))
        # CSR of the Laplacian and of its transpose (~7 non-zeros per row) for the native sparse product on CUDA
        for name, m in (('csr', lap), ('csr_t', lap.T)):
            rows, cols = np.nonzero(m)
            off = np.zeros(self.nv + 1, np.int32)
            np.cumsum(np.bincount(rows, minlength=self.nv), out=off[1:])
            self.register_buffer(name + '_off', torch.from_numpy(off), persistent=False)
            self.register_buffer(name + '_col', torch.from_numpy(cols.astype(np.int32)), persistent=False)
            self.register_buffer(name + '_val', torch.from_numpy(m[rows, cols].astype(np.float32)), persistent=False)

    def forward(self, x):
        if x.is_cuda and x.dim() == 3 and x.dtype == torch.float32:
            y = _LaplacianProduct.apply(x, self).pow(2)
        else:
            y = torch.matmul(self.laplacian, x).pow(2)
        y = y.sum(tuple(range(1, y.dim())))
        return y.sum() / x.size(0) if self.average else y


class _LaplacianProduct(torch.autograd.Function):
    """y = L x with the sparse Laplacian (csrc/scp_geom.cu: scp_spmm3); backward = L^T g."""

    @staticmethod
    def _spmm(mod, t, transpose):
        from ... import _lib
        pre = 'csr_t' if transpose else 'csr'
        off, col, val = (getattr(mod, pre + s) for s in ('_off', '_col', '_val'))
        t = t.contiguous()
        out = torch.empty_like(t)
        with torch.cuda.device(t.device):
            rc = _lib.lib().scp_spmm3(_lib.ptr(off), _lib.ptr(col), _lib.ptr(val), _lib.ptr(t), _lib.ptr(out), t.shape[0],
                                      t.shape[1], _lib.stream_ptr(t.device))
        _lib.check(rc, 'scp_spmm3')
        return out

    @staticmethod
    def forward(ctx, x, mod):
        ctx.mod = mod
        return _LaplacianProduct._spmm(mod, x.detach(), False)

    @staticmethod
    def backward(ctx, g):
        return _LaplacianProduct._spmm(ctx.mod, g, True), None


def compute_mask_loss(img, mask, mask_pred):
    """5-level pyramid of squared differences.  The maps are 3-D (B,H,W), so F.interpolate(mode='area')
    pools along the LAST axis only (rows of the image) -- reproduced as in the reference."""
    total = 0
    for i in range(5):
        d = (F.interpolate(mask_pred, scale_factor=0.5 ** i, mode='area', recompute_scale_factor=False) -
             F.interpolate(mask, scale_factor=0.5 ** i, mode='area', recompute_scale_factor=False)).pow(2)
        total = total + F.interpolate(d[:, None], mask_pred.shape[1:], mode='area')[:, 0]
    return 0.2 * total.mean((1, 2))


def compute_texture_loss(img, mask, tex_pred, tex_mask):
    fg = (mask > 0).float()[:, None]
    img_gt = img * fg
    img_gt_white = 1 - fg + img_gt
    loss = 0.75 * (img_gt - tex_pred * tex_mask[:, None]).pow(2).sum(1).mean((1, 2))
    return loss + (img_gt_white - tex_pred).abs().mean(1).mean((1, 2))


def compute_depth_loss(depth, depth_pred, depth_mask, mask):
    keep = (mask * depth_mask).detach()
    # one batch-global scale; the gradient flows through it (Appendix A.8 of SURVEY.md).  The reference takes
    # the means over boolean-indexed tensors (loss_utils.py:276, a host sync); masked sums give the same value
    # without the data-dependent shape, so the step can be captured in a CUDA graph.
    sel_pred = (depth_mask != 0).to(depth_pred.dtype)
    sel_gt = (mask * depth != 0).to(depth.dtype)
    depth_scale = ((depth_pred * sel_pred).sum() / sel_pred.sum()) / ((depth * sel_gt).sum() / sel_gt.sum())
    diff = depth_pred - depth_scale * depth
    diff = diff.masked_fill((keep == 0) | (depth == 0), 0)
    loss = diff.pow(2)
    loss = 1. - torch.relu(1. - loss)
    return loss.mean((1, 2)), diff


def compute_match_loss(match, match_gt, match_mask, mask):
    valid = (match_mask > 0) & (mask > 0)
    return ((match - match_gt).norm(2, 1) * valid).mean((1, 2))


def compute_imatch_loss(imatch, imatch_gt, depth_weight):
    return ((imatch - imatch_gt).norm(2, 1) * depth_weight).mean(1)


def compute_camera_loss(m1, m2):
    """Geodesic angle between two batches of rotation matrices (loss_utils.py:228-234; flag `camera_loss`, off in every
    shipped config)."""
    m = torch.bmm(m1, m2.transpose(1, 2))
    cos = (m.diagonal(dim1=1, dim2=2).sum(1) - 1) / 2
    return torch.acos(F.hardtanh(cos, -1, 1))


def divide_by_frame(x, batch_size, repeat):
    """(src, tgt) = (frame r, frame r+1 of the same video), wrap-around inside each video."""
    src = x.reshape(batch_size, repeat, *x.shape[1:])
    tgt = torch.roll(src, -1, dims=1)
    return src.reshape(-1, *src.shape[2:]), tgt.reshape(-1, *tgt.shape[2:])


def divide_by_instance(x, batch_size, repeat):
    """(src, tgt) = (video v, same frame slot of video v+1), wrap-around over the batch."""
    src = x.reshape(batch_size, repeat, *x.shape[1:])
    tgt = torch.roll(src, -1, dims=0)
    return src.reshape(-1, *src.shape[2:]), tgt.reshape(-1, *tgt.shape[2:])


def divide_by_both(x, batch_size, repeat):
    sf, tf = divide_by_frame(x, batch_size, repeat)
    si, ti = divide_by_instance(x, batch_size, repeat)
    return torch.cat([sf, si], dim=0), torch.cat([tf, ti], dim=0)
