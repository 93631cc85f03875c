"""CPU: the data-path oracle (oracle/data_cpu.py) against the reference's own dataset class, through the committed golden
vectors (tests/golden/make_data_golden.py ran Wild6DDataset.__getitem__ of /root/reference on a synthetic Wild6D tree)."""
import os

import numpy as np
import pytest
import torch

from oracle import data_cpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'data_golden.npz'))


@pytest.mark.parametrize('no_stretch', [False, True])
@pytest.mark.parametrize('antialias', [False, True])
def test_oracle_getitem_equals_reference_dataset(no_stretch, antialias):
    tag = '%s_%s' % ('nostretch' if no_stretch else 'stretch', 'aa' if antialias else 'noaa')
    n = G['raw_img'].shape[0]
    frames = [(G['raw_img'][i], G['raw_mask'][i], G['raw_depth'][i]) for i in range(n)]
    out = data_cpu.make_batch(frames, list(G['K']), G[tag + '_rand_scale'], 64, no_stretch=no_stretch, antialias=antialias)
    for key in ('mask', 'depth', 'center', 'length', 'foc', 'foc_crop', 'pp', 'pp_crop'):
        assert np.array_equal(out[key].numpy(), G['%s_%s' % (tag, key)]), key      # bit-exact: same statements
    assert np.array_equal(out['img'].float().numpy(), G[tag + '_img'])


def test_golden_covers_the_cases_that_matter():
    lengths = G['stretch_noaa_length']
    crops = 2 * lengths
    assert (crops > 64).any() and (crops < 64).any()              # down- and up-scaling
    c, l = G['stretch_noaa_center'], lengths
    assert ((c - l) < 0).any() or ((c[:, 0] + l[:, 0]) > 320).any() or ((c[:, 1] + l[:, 1]) > 240).any()   # zero padding
    assert np.abs(G['stretch_aa_img'] - G['stretch_noaa_img']).max() > 1e-3   # the two torchvision behaviours differ
