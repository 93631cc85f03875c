"""GPU: the batched crop-box / resized-crop kernels (csrc/scp_data.cu through ops/crop_resize.py and
data/dataset_wild6d.py::GpuBatcher) against the reference's dataset class (golden vectors) and, at the training shape, against
the CPU oracle.  Integer outputs and the nearest-resized maps must be bit-exact; the float64-evaluated bilinear image within
one float32 rounding (1.2e-7) with >= 99.9 % of the pixels bit-identical."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'data_golden.npz'))


def _opts(S, no_stretch=False, use_depth=True):
    return types.SimpleNamespace(img_size=S, no_stretch=no_stretch, use_depth=use_depth)


def _check(out, ref, keys_exact=('mask', 'depth', 'center', 'length', 'foc', 'foc_crop', 'pp', 'pp_crop')):
    for key in keys_exact:
        a, b = out[key].cpu().numpy(), np.asarray(ref[key])
        if key in ('foc_crop', 'pp_crop'):      # float64 products: the device may contract a multiply-add
            assert np.allclose(a, b, rtol=1e-15, atol=1e-12), key
        else:
            assert np.array_equal(a, b), key
    a, b = out['img'].cpu().numpy(), np.asarray(ref['img'], dtype=np.float32)
    d = np.abs(a - b)
    assert d.max() <= 1.2e-7, d.max()
    assert (a == b).mean() >= 0.999, (a == b).mean()
    return d.max(), (a == b).mean()


@pytest.mark.parametrize('no_stretch', [False, True])
@pytest.mark.parametrize('antialias', [False, True])
def test_batch_equals_reference_dataset(no_stretch, antialias):
    from self_corr_pose_b200.data.dataset_wild6d import GpuBatcher
    tag = '%s_%s' % ('nostretch' if no_stretch else 'stretch', 'aa' if antialias else 'noaa')
    n = G['raw_img'].shape[0]
    frames = [(G['raw_img'][i], G['raw_mask'][i], G['raw_depth'][i]) for i in range(n)]
    out = GpuBatcher(_opts(64, no_stretch), antialias=antialias).make_batch(frames, list(G['K']), list(range(n)),
                                                                            rand_scale=G[tag + '_rand_scale'])
    assert int(out['status'].sum()) == 0
    ref = {k: G['%s_%s' % (tag, k)] for k in ('img', 'mask', 'depth', 'center', 'length', 'foc', 'foc_crop', 'pp', 'pp_crop')}
    mx, same = _check(out, ref)
    print('PARITY data %s max|d| %.1e identical %.5f' % (tag, mx, same))


@pytest.mark.parametrize('antialias', [False, True])
def test_training_shape_vs_oracle(antialias):
    """640 x 480 frames -> 256 x 256, B = 8 (crop boxes from 50 to 670 pixels, several leaving the frame)."""
    from oracle import data_cpu
    from self_corr_pose_b200.data.dataset_wild6d import GpuBatcher
    frames, Ks = data_cpu.synthetic_frames(8, 480, 640, seed=3)
    rs = np.random.RandomState(4).uniform(1.2, 1.5, size=(8, 2))
    ref = data_cpu.make_batch(frames, Ks, rs, 256, antialias=antialias)
    ref = {k: v.numpy() for k, v in ref.items()}
    out = GpuBatcher(_opts(256), antialias=antialias).make_batch(frames, Ks, list(range(8)), rand_scale=rs)
    mx, same = _check(out, ref)
    print('PARITY data 640x480->256 aa=%d max|d| %.1e identical %.5f' % (antialias, mx, same))


def test_empty_mask_is_flagged_and_rgb_order():
    from self_corr_pose_b200.ops import crop_resize
    mask = torch.zeros(2, 32, 48, dtype=torch.uint8, device='cuda')
    mask[1, 10:20, 5:30] = 255
    box = crop_resize.bbox_crop(mask, torch.full((2, 2), 1.25, dtype=torch.float64), torch.ones(2, 4, dtype=torch.float64), 16)
    assert box['status'].tolist() == [1, 0]
    assert box['crop'][0].tolist() == [0, 0, 0, 0]
    assert box['center'][1].tolist() == [(29 + 5) // 2, (19 + 10) // 2] and box['length'][1].tolist() == [int(1.25 * 12), int(1.25 * 4)]
    img = torch.zeros(2, 32, 48, 3, dtype=torch.uint8, device='cuda')
    img[..., 0] = 255                     # first stored channel
    rgb, m, _ = crop_resize.resized_crop(img, mask, None, box['crop'], 16, bgr=False)
    bgr, _, _ = crop_resize.resized_crop(img, mask, None, box['crop'], 16, bgr=True)
    assert float(rgb[0].abs().max()) == 0 and float(m[0].abs().max()) == 0       # empty crop -> zeros
    inside = m[1, 0] > 0
    assert float(rgb[1, 0][inside].min()) == 1.0 and float(bgr[1, 2][inside].min()) == 1.0 and float(bgr[1, 0].max()) == 0.0


def test_batch_feeds_the_trainer_layout():
    """The dict has the reference loader's keys / shapes / dtypes (what Trainer.batch_reshape consumes)."""
    from oracle import data_cpu
    from self_corr_pose_b200.data.dataset_wild6d import GpuBatcher
    frames, Ks = data_cpu.synthetic_frames(4, 120, 160, seed=1)
    np.random.seed(0)
    out = GpuBatcher(_opts(64)).make_batch(frames, Ks, [0, 0, 1, 1])
    assert out['img'].shape == (4, 3, 64, 64) and out['img'].dtype == torch.float32
    assert out['mask'].shape == (4, 1, 64, 64) and out['depth'].shape == (4, 1, 64, 64)
    assert out['foc_crop'].dtype == torch.float64 and out['pp_crop'].shape == (4, 2)
    assert out['center'].dtype == torch.int64 and out['idx'].shape == (4, 1)
    np.random.seed(0)
    rs = np.stack([np.random.uniform(1.2, 1.5, size=(2,)) for _ in range(4)])       # same draws as the reference loop
    ref = data_cpu.make_batch(frames, Ks, rs, 64)
    assert np.array_equal(out['length'].cpu().numpy(), ref['length'].numpy())
