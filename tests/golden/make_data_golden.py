"""Generates tests/golden/data_golden.npz by running the REFERENCE's own training dataset class
(data/dataset_wild6d.py: Wild6DDataset.__getitem__) on a small synthetic Wild6D directory written here with cv2
(JPEG frames, PNG masks, 16-bit PNG depth, a `metadata` JSON per sequence), in the build container.
Stored per sample: the decoded files exactly as the reference's cv2.imread calls return them (the kernel's inputs), the
rand_scale draw, and every entry of the returned dict -- once as the reference runs under this container's torchvision
(0.26: `resized_crop` applies the antialias filter to tensors by default) and once with torchvision 0.11's behaviour, the
reference's pinned environment (README.md:25), obtained by passing antialias=False through the same call.
Two option sets: stretch (laptop config) and no_stretch.
Run:  python tests/golden/make_data_golden.py      (needs /root/reference; the .npz is committed)
"""
import json
import os
import sys
import tempfile
import types

import cv2
import numpy as np
import torch

REF = '/root/reference'
sys.path.insert(0, REF)
sys.modules.setdefault('tqdm', types.ModuleType('tqdm'))
if not hasattr(sys.modules['tqdm'], 'tqdm'):
    sys.modules['tqdm'].tqdm = lambda x, *a, **k: x
from data.dataset_wild6d import Wild6DDataset   # noqa: E402
from torchvision import transforms              # noqa: E402

H, W, S = 240, 320, 64


def write_dataset(root):
    rng = np.random.RandomState(0)
    # two objects x one sequence, 4 frames each; ellipse silhouettes of different sizes (crop smaller and larger than S),
    # one touching the frame border (crop box leaves the frame -> zero padding)
    shapes = [[(160, 120, 22, 15), (100, 90, 60, 40), (30, 200, 28, 35), (250, 60, 90, 55)],
              [(300, 20, 30, 25), (150, 130, 110, 80), (80, 60, 12, 9), (200, 200, 45, 38)]]
    for oi, frames in enumerate(shapes):
        d = os.path.join(root, 'obj%d' % oi, 'seq0', 'images')
        os.makedirs(d)
        K = [300.0 + 10 * oi, 0, 0, 0, 310.0 + 5 * oi, 0, W / 2 + 3.5, H / 2 - 2.25, 1]     # stored transposed (column-major)
        json.dump({'K': K, 'w': W, 'h': H, 'fps': 30}, open(os.path.join(root, 'obj%d' % oi, 'seq0', 'metadata'), 'w'))
        yy, xx = np.mgrid[0:H, 0:W]
        for fi, (cx, cy, ax, ay) in enumerate(frames):
            m = (((xx - cx) / ax) ** 2 + ((yy - cy) / ay) ** 2 <= 1.0)
            img = (rng.rand(H, W, 3) * 60 + 40 * np.sin(xx / 9.0 + fi)[..., None] + 90).clip(0, 255)
            img[m] = (img[m] * 0.5 + np.array([200, 120, 60]) * 0.5)
            depth = (900 + 3 * xx + 2 * yy + 40 * fi).astype(np.uint16) * m
            cv2.imwrite(os.path.join(d, '%d.jpg' % fi), img.astype(np.uint8))
            cv2.imwrite(os.path.join(d, '%d-mask.png' % fi), (m * 255).astype(np.uint8))
            cv2.imwrite(os.path.join(d, '%d-depth.png' % fi), depth)
    open(os.path.join(root, 'train_list.txt'), 'w').write('laptop_0_0\nlaptop_1_0\n')


def run(root, no_stretch, antialias):
    opts = types.SimpleNamespace(train_list=os.path.join(root, 'train_list.txt'), dataset_path=root, batch_size=2, repeat=2,
                                 ngpu=1, total_iters=2, use_depth=True, no_stretch=no_stretch, img_size=S)
    np.random.seed(5)
    ds = Wild6DDataset(opts)
    orig = transforms.functional.resized_crop
    if not antialias:      # torchvision 0.11 semantics through the reference's own call
        transforms.functional.resized_crop = lambda *a, **k: orig(*a, **dict(k, antialias=False))
    out = []
    try:
        np.random.seed(11)
        for index in range(len(ds)):
            state = np.random.get_state()
            rs = np.random.uniform(1.2, 1.5, size=(2,))      # what __getitem__ draws first
            np.random.set_state(state)
            elem = ds[index]
            vid, fid = int(elem['idx']), int(elem['frame_idx'])
            raw = (cv2.imread(ds.imglist[vid][fid]), cv2.imread(ds.masklist[vid][fid], cv2.IMREAD_GRAYSCALE),
                   cv2.imread(ds.depthlist[vid][fid], -1))
            out.append((rs, raw, ds.metalist[vid][0], elem))
    finally:
        transforms.functional.resized_crop = orig
    return out


def main():
    root = tempfile.mkdtemp()
    write_dataset(root)
    gold = {}
    for no_stretch in (False, True):
        for antialias in (True, False):
            tag = '%s_%s' % ('nostretch' if no_stretch else 'stretch', 'aa' if antialias else 'noaa')
            samples = run(root, no_stretch, antialias)
            gold[tag + '_rand_scale'] = np.stack([s[0] for s in samples])
            raw = dict(raw_img=np.stack([s[1][0] for s in samples]), raw_mask=np.stack([s[1][1] for s in samples]),
                       raw_depth=np.stack([s[1][2] for s in samples]), K=np.stack([s[2] for s in samples]))
            for k, v in raw.items():      # the same frames are sampled for every option set (same seeds): stored once
                assert k not in gold or np.array_equal(gold[k], v)
                gold[k] = v
            for key in ('img', 'mask', 'depth', 'center', 'length', 'foc', 'foc_crop', 'pp', 'pp_crop', 'idx', 'frame_idx'):
                v = torch.stack([s[3][key] for s in samples])
                if key == 'img':
                    assert v.dtype == torch.float64
                    v = v.float()          # Trainer.batch_reshape: batch['img'].float() (model/trainer.py:82)
                gold['%s_%s' % (tag, key)] = v.numpy()
            print(tag, 'img dtype', samples[0][3]['img'].dtype, 'crops (2*length):',
                  [tuple(int(2 * x) for x in s[3]['length']) for s in samples])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data_golden.npz')
    np.savez_compressed(path, **gold)
    print(path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
