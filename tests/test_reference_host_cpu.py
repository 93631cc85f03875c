"""CPU: host-side classes of the training step against the REFERENCE's own modules imported from /root/reference
(skipped where the reference tree is not mounted, e.g. on the GPU box): loss-weight schedule (model/module/weights.py)
and the optimiser's parameter grouping + OneCycle learning-rate trace (model/module/optimizers.py)."""
import importlib.util
import os
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'model')), reason='reference tree not mounted')


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _opts():
    from self_corr_pose_b200.hotpath import default_opts
    return default_opts(total_iters=200, ngpu=1)


def test_weight_schedule_matches_reference():
    from self_corr_pose_b200.model.module.weights import Weights
    ref = _load('model/module/weights.py', 'ref_weights')
    opts = _opts()
    a, b = Weights(opts), ref.Weights(opts)
    names = ('mask_wt', 'depth_wt', 'tex_wt', 'match_wt', 'imatch_wt', 'triangle_wt', 'pullfar_wt', 'deform_wt',
             'symmetry_wt', 'camera_wt', 'cycle_loss_wt', 'cycle_loss_pt_wt')
    for it in (0, 1, 37, 100, 199, 200, 201, 1000):
        a.schedule(it)
        b.schedule(it)
        for n in names:
            assert getattr(a, n) == getattr(b, n), (it, n)


class _Toy(nn.Module):
    """parameter names that exercise every branch of the reference's name-keyed grouping"""

    def __init__(self):
        super().__init__()
        self.mesh = nn.ParameterDict({'mean_v': nn.Parameter(torch.randn(5, 3)),
                                      'faces': nn.Parameter(torch.zeros(4, 3), requires_grad=False)})
        self.encoder = nn.ModuleDict({
            'pose_predictor': nn.Linear(4, 3), 'shape_predictor': nn.Linear(4, 2), 'shape_code_predictor': nn.Linear(4, 2),
            'featnet': nn.Linear(4, 4), 'backbone': nn.Linear(4, 4), 'mesh_featnet': nn.Linear(3, 4)})
        self.pretrain_corr_net = nn.Linear(2, 2)
        self.other = nn.Linear(2, 2)


def test_optimizer_groups_and_lr_trace_match_reference(capsys):
    from self_corr_pose_b200.model.module.optimizers import Optimizers
    ref = _load('model/module/optimizers.py', 'ref_optimizers')
    opts = _opts()
    torch.manual_seed(0)
    m1, m2 = _Toy(), _Toy()
    m2.load_state_dict(m1.state_dict())
    a, b = Optimizers(opts, m1), ref.Optimizers(opts, m2)
    capsys.readouterr()                                    # the reference prints every parameter it finds
    ga = [[tuple(p.shape) for p in g['params']] for g in a.optimizer.param_groups]
    gb = [[tuple(p.shape) for p in g['params']] for g in b.optimizer.param_groups]
    assert ga == gb
    x = torch.randn(3, 4)
    for it in range(60):
        for m, o in ((m1, a), (m2, b)):
            o.zero_grad()
            loss = sum(v(x).pow(2).mean() for k, v in m.encoder.items() if k != 'mesh_featnet') + m.mesh['mean_v'].pow(2).sum()
            loss.backward()
            o.step(it)
        assert [g['lr'] for g in a.optimizer.param_groups] == [g['lr'] for g in b.optimizer.param_groups]
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(p, q)
