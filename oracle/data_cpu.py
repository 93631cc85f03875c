"""TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline): CPU restatement of the per-frame work of the
reference's training loader after the file decode -- Wild6DDataset.__getitem__, data/dataset_wild6d.py:122-163 -- on decoded
frames: numpy bounding box of the mask (:131-137), crop box and crop intrinsics (:138-147), ToTensor / 255 on the float64 image
(:149) and the three torchvision `resized_crop` calls (:152-160), statement by statement.  Pinned against the reference class
itself through tests/golden/data_golden.npz (tests/test_oracle_data.py).  `antialias` selects what torchvision >= 0.17 does by
default for tensors (True) or torchvision 0.11, the reference's pinned environment (False)."""
import numpy as np
import torch
from torchvision import transforms
from torchvision.transforms import InterpolationMode


def getitem(img_bgr, mask_u8, depth_u16, K, rand_scale, img_size, no_stretch=False, use_depth=True, antialias=False):
    img = img_bgr[:, :, ::-1]
    mask = mask_u8.astype(bool)
    if use_depth:
        depth = depth_u16 * 1.0
    img = img * 1.0
    indices = np.where(mask > 0)
    xid, yid = indices[1], indices[0]
    center = [(xid.max() + xid.min()) // 2, (yid.max() + yid.min()) // 2]
    length = [(xid.max() - xid.min()) // 2, (yid.max() - yid.min()) // 2]
    max_length = max(length[0], length[1])
    if no_stretch:
        length = [int(rand_scale[0] * max_length), int(rand_scale[0] * max_length)]
    else:
        length = [int(rand_scale[0] * length[0]), int(rand_scale[1] * length[1])]
    foc = [K[0, 0], K[1, 1]]
    pp = [K[0, 2], K[1, 2]]
    maxw = maxh = img_size
    crop_factor = [maxw / 2 / length[0], maxh / 2 / length[1]]
    foc_crop = [foc[0] * crop_factor[0], foc[1] * crop_factor[1]]
    pp_crop = [(pp[0] - (center[0] - length[0])) * crop_factor[0], (pp[1] - (center[1] - length[1])) * crop_factor[1]]
    img = transforms.ToTensor()(img) / 255.
    mask = torch.tensor(mask, dtype=torch.float32)[None]
    box = (center[1] - length[1], center[0] - length[0], 2 * length[1], 2 * length[0])
    img = transforms.functional.resized_crop(img, *box, size=(maxh, maxw), interpolation=InterpolationMode.BILINEAR,
                                             antialias=antialias)
    mask = transforms.functional.resized_crop(mask, *box, size=(maxh, maxw), interpolation=InterpolationMode.NEAREST)
    if use_depth:
        depth = torch.tensor(depth, dtype=torch.float32)[None]
        depth = transforms.functional.resized_crop(depth, *box, size=(maxh, maxw), interpolation=InterpolationMode.NEAREST)
    return {'img': img, 'mask': mask, 'depth': depth if use_depth else torch.zeros(1), 'center': torch.tensor(center),
            'length': torch.tensor(length), 'foc': torch.tensor(foc), 'foc_crop': torch.tensor(foc_crop),
            'pp': torch.tensor(pp), 'pp_crop': torch.tensor(pp_crop)}


def make_batch(frames, Ks, rand_scale, img_size, **kw):
    elems = [getitem(f[0], f[1], f[2], K, rs, img_size, **kw) for f, K, rs in zip(frames, Ks, rand_scale)]
    return {k: torch.stack([e[k] for e in elems]) for k in elems[0]}


def synthetic_frames(B, H, W, seed=0):
    """Decoded-frame stand-ins: noisy BGR image, one ellipse silhouette per frame (some leaving the frame), 16-bit depth."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    frames, Ks = [], []
    for b in range(B):
        cx, cy = rng.uniform(0.1, 0.9) * W, rng.uniform(0.1, 0.9) * H
        ax, ay = rng.uniform(0.04, 0.35) * W, rng.uniform(0.04, 0.35) * H
        m = (((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0)
        if not m.any():
            m[H // 2, W // 2] = True
        img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
        depth = (rng.randint(400, 3000, size=(H, W)) * m).astype(np.uint16)
        frames.append((img, (m * 255).astype(np.uint8), depth))
        Ks.append(np.array([[rng.uniform(400, 700), 0, W / 2 + rng.uniform(-9, 9)], [0, rng.uniform(400, 700), H / 2 + rng.uniform(-9, 9)],
                            [0, 0, 1]]))
    return frames, Ks
