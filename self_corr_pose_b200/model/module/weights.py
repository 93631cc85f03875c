"""Loss weights and their linear schedule (reference: model/module/weights.py:20-64)."""
import numpy as np


def reg_decay(curr_steps, max_steps, min_wt, max_wt, mode='linear'):
    if curr_steps > max_steps:
        return min_wt
    if mode == 'log':
        return np.exp(curr_steps / float(max_steps) * (np.log(min_wt) - np.log(max_wt))) * max_wt
    if mode == 'linear':
        return curr_steps / float(max_steps) * (min_wt - max_wt) + max_wt
    raise NotImplementedError


class Weights:
    _names = ('mask_wt', 'depth_wt', 'tex_wt', 'match_wt', 'imatch_wt', 'triangle_wt', 'pullfar_wt', 'deform_wt',
              'symmetry_wt', 'camera_wt', 'cycle_loss_wt')

    def __init__(self, opts):
        self.opts = opts
        self.total_iters = opts.total_iters
        for n in self._names:
            setattr(self, n, getattr(opts, n))
        self.cycle_loss_pt_wt = opts.cycle_loss_pretrain_wt

    def schedule(self, it):
        o, T = self.opts, self.total_iters
        # decreasing regularisers / cycle terms, increasing correspondence terms
        self.triangle_wt = reg_decay(it, T, o.decay_ratio * o.triangle_wt, o.triangle_wt)
        self.symmetry_wt = reg_decay(it, T, o.decay_ratio * o.symmetry_wt, o.symmetry_wt)
        self.cycle_loss_wt = reg_decay(it, T, o.decay_ratio * o.cycle_loss_wt, o.cycle_loss_wt)
        self.cycle_loss_pt_wt = reg_decay(it, T, o.decay_ratio * o.cycle_loss_pretrain_wt, o.cycle_loss_pretrain_wt)
        self.match_wt = reg_decay(it, T, o.match_wt, o.decay_ratio * o.match_wt)
        self.imatch_wt = reg_decay(it, T, o.imatch_wt, o.decay_ratio * o.imatch_wt)
