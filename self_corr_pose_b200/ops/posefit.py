"""The two point-cloud passes of the batched RANSAC + Umeyama pose fit on the sm_100a kernels (csrc/scp_posefit.cu;
SURVEY.md section 8f-2, model/util/umeyama.py:95-159 of the reference): the residual table of all candidate transforms of
all images, and the inlier set / closed-form moments of every image's winning round.  CUDA tensors only."""
import torch

from .. import _lib


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def residual_table(src, tgt, counts_dev, hyp_s, hyp_R, hyp_t):
    """src, tgt (L, n, 3); counts_dev (L,) int32 real points per image; candidates hyp_s (L, H), hyp_R (L, H, 3, 3),
    hyp_t (L, H, 3) -> (L, H) residual norms || tgt - (s R src + t) || over each image's real points (evaluateModel of
    the reference, umeyama.py:143-159, for every round at once)."""
    if not src.is_cuda:
        raise TypeError('posefit.residual_table supports only CUDA tensors (the host formulation lives in model/util/umeyama.py)')
    L, n, _ = src.shape
    H = hyp_s.shape[1]
    _lib.expect_numel('posefit.residual_table', tgt=(tgt, L * n * 3), counts=(counts_dev, L), hyp_R=(hyp_R, L * H * 9),
                      hyp_t=(hyp_t, L * H * 3))
    # every converted operand stays referenced until the launch has been enqueued
    src, tgt, A, hyp_t = _f32(src), _f32(tgt), _f32(hyp_s[..., None, None] * hyp_R), _f32(hyp_t)
    lib = _lib.lib()
    nchunk = lib.scp_posefit_chunks(n)
    partial = torch.empty(L, nchunk, H, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        rc = lib.scp_posefit_residual_table(_lib.ptr(src), _lib.ptr(tgt), _lib.ptr(counts_dev), _lib.ptr(A), _lib.ptr(hyp_t),
                                            L, n, H, _lib.ptr(partial), _lib.stream_ptr(src.device))
    _lib.check(rc, 'scp_posefit_residual_table')
    return partial.sum(1).sqrt()


def inlier_moments(src, tgt, counts_dev, best_s, best_R, best_t, pass_t, found):
    """Winning transform of every image -> dict(n_used, n_inliers, mean_src (L,3), mean_tgt (L,3), cov (L,3,3) = centred
    sum of tgt src^T, sq (L,) = centred sum of |src|^2) over the inliers (images with an accepted round and at least 10 %
    inliers) or over all real points (rejected images: keeps the closed form finite, the caller discards it)."""
    if not src.is_cuda:
        raise TypeError('posefit.inlier_moments supports only CUDA tensors')
    L, n, _ = src.shape
    _lib.expect_numel('posefit.inlier_moments', tgt=(tgt, L * n * 3), counts=(counts_dev, L), best_R=(best_R, L * 9),
                      best_t=(best_t, L * 3), pass_t=(pass_t, L), found=(found, L))
    src, tgt, A, best_t, pass_t = _f32(src), _f32(tgt), _f32(best_s[..., None, None] * best_R), _f32(best_t), _f32(pass_t)
    found = found.to(torch.uint8).contiguous()
    out = torch.empty(L, 18, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        rc = _lib.lib().scp_posefit_inlier_moments(_lib.ptr(src), _lib.ptr(tgt), _lib.ptr(counts_dev), _lib.ptr(A),
                                                   _lib.ptr(best_t), _lib.ptr(pass_t), _lib.ptr(found), L, n, _lib.ptr(out),
                                                   _lib.stream_ptr(src.device))
    _lib.check(rc, 'scp_posefit_inlier_moments')
    return dict(n_used=out[:, 0], n_inliers=out[:, 1], mean_src=out[:, 2:5], mean_tgt=out[:, 5:8],
                cov=out[:, 8:17].reshape(L, 3, 3), sq=out[:, 17])
