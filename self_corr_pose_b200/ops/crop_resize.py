"""Batched crop box + resized crop of decoded frames (csrc/scp_data.cu): the per-frame work of the reference's
Wild6DDataset.__getitem__ (data/dataset_wild6d.py:131-163) after the files are decoded, on the GPU.  No CPU path."""
import torch

from .. import _lib


def bbox_crop(mask, rand_scale, intr, img_size, no_stretch=False):
    """mask (B,H,W) uint8 CUDA (non-zero = foreground); rand_scale (B,2) float64; intr (B,4) float64 = fx, fy, cx, cy.
    Returns dict(crop (B,4) int32 = top, left, height, width; center, length (B,2) int64; foc_crop, pp_crop (B,2) float64;
    status (B,) int32, 1 = empty silhouette) -- all on the device, no host synchronisation."""
    if not mask.is_cuda:
        raise TypeError('bbox_crop supports only CUDA tensors (no CPU path)')
    assert mask.dtype == torch.uint8 and mask.dim() == 3
    dev = mask.device
    mask = mask.contiguous()
    B, H, W = mask.shape
    rand_scale = rand_scale.to(dev, torch.float64).contiguous()
    intr = intr.to(dev, torch.float64).contiguous()
    assert rand_scale.shape == (B, 2) and intr.shape == (B, 4)
    out = dict(crop=torch.empty(B, 4, dtype=torch.int32, device=dev), center=torch.empty(B, 2, dtype=torch.int64, device=dev),
               length=torch.empty(B, 2, dtype=torch.int64, device=dev), foc_crop=torch.empty(B, 2, dtype=torch.float64, device=dev),
               pp_crop=torch.empty(B, 2, dtype=torch.float64, device=dev), status=torch.empty(B, dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        rc = _lib.lib().scp_data_bbox_crop(_lib.ptr(mask), _lib.ptr(rand_scale), _lib.ptr(intr), B, H, W, int(img_size),
                                           1 if no_stretch else 0, _lib.ptr(out['crop']), _lib.ptr(out['center']),
                                           _lib.ptr(out['length']), _lib.ptr(out['foc_crop']), _lib.ptr(out['pp_crop']),
                                           _lib.ptr(out['status']), _lib.stream_ptr(dev))
    _lib.check(rc, 'scp_data_bbox_crop')
    return out


def resized_crop(img, mask, depth, crop, img_size, bgr=False, antialias=False, out=None):
    """img (B,H,W,3) uint8, mask (B,H,W) uint8, depth (B,H,W) uint16 (or int16 storage) or None, crop (B,4) int32 -- CUDA.
    Returns (img (B,3,S,S) fp32 in [0,1], mask (B,1,S,S) fp32, depth (B,1,S,S) fp32 or None), the layouts of the reference's
    batch; `out` = (img, mask, depth) buffers to write into (e.g. the static batch of Trainer.capture)."""
    if not img.is_cuda:
        raise TypeError('resized_crop supports only CUDA tensors (no CPU path)')
    dev = img.device
    img, mask = img.contiguous(), mask.contiguous()
    B, H, W, _ = img.shape
    S = int(img_size)
    assert img.dtype == torch.uint8 and mask.dtype == torch.uint8 and crop.dtype == torch.int32
    if depth is not None:
        assert depth.dtype in (torch.uint16, torch.int16) and depth.shape == (B, H, W)
        depth = depth.contiguous()
    if out is None:
        out = (torch.empty(B, 3, S, S, device=dev), torch.empty(B, 1, S, S, device=dev),
               torch.empty(B, 1, S, S, device=dev) if depth is not None else None)
    o_img, o_mask, o_depth = out
    with torch.cuda.device(dev):
        rc = _lib.lib().scp_data_resized_crop(_lib.ptr(img), _lib.ptr(mask), _lib.ptr(depth) if depth is not None else None,
                                              _lib.ptr(crop.contiguous()), B, H, W, S, 1 if bgr else 0, 1 if antialias else 0,
                                              _lib.ptr(o_img), _lib.ptr(o_mask), _lib.ptr(o_depth) if depth is not None else None,
                                              _lib.stream_ptr(dev))
    _lib.check(rc, 'scp_data_resized_crop')
    return o_img, o_mask, o_depth
