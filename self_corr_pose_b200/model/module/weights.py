"""Loss weights of the training step and their schedule over the iterations.

Behaviour of the reference's model/module/weights.py:20-64 (pinned by tests/test_reference_host_cpu.py): the
regularisers and cycle terms decay linearly from their flag value to `decay_ratio` times it over `total_iters`, the two
correspondence terms grow the other way round, everything else is constant."""
import math

from ... import _flags

_CONSTANT = ('mask_wt', 'depth_wt', 'tex_wt', 'pullfar_wt', 'deform_wt', 'camera_wt')
# attribute -> (flag holding the base value, True = starts at the base value and decays to decay_ratio * base)
_SCHEDULED = {
    'triangle_wt': ('triangle_wt', True),
    'symmetry_wt': ('symmetry_wt', True),
    'cycle_loss_wt': ('cycle_loss_wt', True),
    'cycle_loss_pt_wt': ('cycle_loss_pretrain_wt', True),
    'match_wt': ('match_wt', False),
    'imatch_wt': ('imatch_wt', False),
}


def reg_decay(curr_steps, max_steps, min_wt, max_wt, mode='linear'):
    """Value at `curr_steps` of a ramp from max_wt (step 0) to min_wt (step max_steps), constant afterwards."""
    if curr_steps > max_steps:
        return min_wt
    frac = curr_steps / float(max_steps)
    if mode == 'linear':
        return frac * (min_wt - max_wt) + max_wt
    if mode == 'log':
        return math.exp(frac * (math.log(min_wt) - math.log(max_wt))) * max_wt
    raise NotImplementedError(mode)


class Weights:

    def __init__(self, opts):
        self.opts = opts
        self.total_iters = opts.total_iters
        for name in _CONSTANT:
            setattr(self, name, getattr(opts, name))
        for name, (flag, _) in _SCHEDULED.items():
            setattr(self, name, getattr(opts, flag))

        self._dev = None

    def values(self, it):
        """{name: weight at iteration `it`} -- constants and scheduled ones."""
        ratio = self.opts.decay_ratio
        out = {name: getattr(self.opts, name) for name in _CONSTANT}
        for name, (flag, decays) in _SCHEDULED.items():
            base = getattr(self.opts, flag)
            end_points = (ratio * base, base) if decays else (base, ratio * base)     # (value at the end, value at step 0)
            out[name] = reg_decay(it, self.total_iters, *end_points)
        return out

    def schedule(self, it):
        vals = self.values(it)
        if self._dev is None:
            for name in _SCHEDULED:
                setattr(self, name, vals[name])
            return
        # static mode: the weights are 0-dim views of one device buffer, refreshed through a (capturable) copy from pinned
        # host memory -- a CUDA graph of the step then follows the schedule: fill_host(it) before every replay
        self.fill_host(it)
        self._dev.copy_(self._host, non_blocking=True)

    def enable_device_buffer(self, device):
        """Switches the weights from python floats to device scalars (same values: a float32 factor either way)."""
        import torch
        self._names = list(_CONSTANT) + list(_SCHEDULED)
        self._host = _flags.pinned(torch.zeros(len(self._names), dtype=torch.float32))
        self._dev = torch.zeros(len(self._names), dtype=torch.float32, device=device)
        for i, name in enumerate(self._names):
            setattr(self, name, self._dev[i])

    def fill_host(self, it):
        vals = self.values(it)
        for i, name in enumerate(self._names):
            self._host[i] = float(vals[name])
