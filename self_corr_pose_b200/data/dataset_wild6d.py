"""GPU side of the training loader.  The reference's Wild6DDataset.__getitem__
(data/dataset_wild6d.py:115-182) decodes three files per frame with cv2 and then crops / resizes them one frame at a time on
a DataLoader worker; at the step rates of this package (>= 1 400 images/s per GPU) eight workers cannot keep up.  Here the
workers only decode (`decode_frame`, the reference's three cv2.imread calls); a batch of decoded uint8 / uint16 frames is
uploaded once and `GpuBatcher.make_batch` produces the reference's batch dict with two kernel launches (ops/crop_resize.py)
-- optionally straight into the static input buffers of a captured training step.

Random numbers: `rand_scale = np.random.uniform(1.2, 1.5, size=(2,))` is drawn per frame, in batch order, from the global
numpy generator exactly as the reference's __getitem__ does (one worker process, in-order).

antialias: the reference's pinned torchvision 0.11 resizes tensors WITHOUT the antialias filter; torchvision >= 0.17 applies it
by default to the same `resized_crop` call.  Both are implemented; the default follows the reference's environment."""
import numpy as np
import torch

from ..ops import crop_resize


def decode_frame(img_path, mask_path, depth_path=None):
    """The file reads of dataset_wild6d.py:125-129 (BGR image as cv2 returns it, grey mask, 16-bit depth)."""
    import cv2
    img = cv2.imread(img_path)
    mask = cv2.imread(mask_path, cv2.IMREAD_GRAYSCALE)
    depth = cv2.imread(depth_path, -1) if depth_path is not None else None
    return img, mask, depth


class GpuBatcher:
    def __init__(self, opts, device='cuda', antialias=False):
        self.opts, self.device, self.antialias = opts, torch.device(device), antialias

    def make_batch(self, frames, intrinsics, video_ids, frame_ids=None, rand_scale=None, out=None, bgr=True):
        """frames: list of (img (H,W,3) uint8 BGR, mask (H,W) uint8, depth (H,W) uint16 or None) numpy arrays of ONE size;
        intrinsics: list of 3x3 K (as metalist holds them); returns the dict DataLoader would collate from the reference's
        __getitem__ elems (img, mask, depth, center, length, foc, foc_crop, pp, pp_crop, idx, frame_idx), tensors on the GPU."""
        opts, dev = self.opts, self.device
        B = len(frames)
        if rand_scale is None:
            rand_scale = np.stack([np.random.uniform(1.2, 1.5, size=(2,)) for _ in range(B)])
        use_depth = bool(getattr(opts, 'use_depth', True))
        img = torch.from_numpy(np.stack([f[0] for f in frames])).to(dev, non_blocking=True)
        mask = torch.from_numpy(np.stack([f[1] for f in frames])).to(dev, non_blocking=True)
        depth = None
        if use_depth:
            depth = torch.from_numpy(np.stack([f[2] for f in frames]).astype(np.uint16).view(np.int16)).to(dev, non_blocking=True)
        K = np.stack([np.asarray(k, dtype=np.float64) for k in intrinsics])
        intr = torch.from_numpy(np.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]], axis=1))
        return self.make_batch_device(img, mask, depth, intr, torch.from_numpy(np.asarray(rand_scale, dtype=np.float64)),
                                      video_ids, frame_ids, out=out, bgr=bgr)

    def make_batch_device(self, img, mask, depth, intr, rand_scale, video_ids, frame_ids=None, out=None, bgr=True):
        """Same with the decoded frames already in HBM: img (B,H,W,3) uint8, mask (B,H,W) uint8, depth (B,H,W) int16 storage
        of the uint16 values, intr (B,4) float64 = fx, fy, cx, cy."""
        opts, dev = self.opts, self.device
        S = int(opts.img_size)
        B = img.shape[0]
        box = crop_resize.bbox_crop(mask, rand_scale, intr, S, bool(getattr(opts, 'no_stretch', False)))
        o_img, o_mask, o_depth = crop_resize.resized_crop(img, mask, depth, box['crop'], S, bgr=bgr, antialias=self.antialias,
                                                          out=out)
        intr = intr.to(dev)
        elem = {'img': o_img, 'mask': o_mask,
                'depth': o_depth if depth is not None else torch.zeros(B, 1, device=dev),
                'center': box['center'], 'length': box['length'],
                'foc': intr[:, 0:2].contiguous(), 'foc_crop': box['foc_crop'],
                'pp': intr[:, 2:4].contiguous(), 'pp_crop': box['pp_crop'],
                'idx': torch.as_tensor(video_ids, dtype=torch.int64, device=dev).reshape(B, 1),
                'frame_idx': torch.as_tensor(frame_ids if frame_ids is not None else [0] * B, dtype=torch.int64,
                                             device=dev).reshape(B, 1),
                'status': box['status']}
        return elem
