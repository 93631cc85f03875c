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


def _reference_trainer_methods():
    """batch_reshape / collect_grad of the reference's Trainer, compiled from its source without importing the module
    (which pulls in the dataset, TensorBoard and the CUDA-only model)."""
    import ast
    import textwrap
    src = open(os.path.join(REF, 'model', 'trainer.py')).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'Trainer')
    ns = {'torch': torch}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ('batch_reshape', 'collect_grad'):
            exec(compile(ast.Module(body=[fn], type_ignores=[]), 'reference_trainer', 'exec'), ns)
    return ns['batch_reshape'], ns['collect_grad']


def test_trainer_batch_reshape_and_collect_grad_match_reference(monkeypatch, capsys):
    from self_corr_pose_b200.model.trainer import Trainer
    ref_reshape, ref_collect = _reference_trainer_methods()
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    opts = _opts()
    g = torch.Generator().manual_seed(3)
    B = 4
    batch = {'img': torch.rand(B, 3, 16, 16, generator=g).double(), 'mask': (torch.rand(B, 1, 16, 16, generator=g) > 0.5).float(),
             'depth': torch.rand(B, 1, 16, 16, generator=g), 'center': torch.zeros(B, 2, dtype=torch.int64),
             'length': torch.full((B, 2), 256, dtype=torch.int64), 'idx': torch.arange(B)[:, None],
             'foc': torch.rand(B, 2, generator=g).double() * 500, 'pp': torch.rand(B, 2, generator=g).double() * 256,
             'foc_crop': torch.rand(B, 2, generator=g).double() * 500, 'pp_crop': torch.rand(B, 2, generator=g).double() * 256}
    t = Trainer(opts)
    t.device = torch.device('cpu')
    got = t.batch_reshape(batch)
    want = ref_reshape(SimpleNamespace(opts=opts), batch)
    assert len(got) == len(want) == 12
    for a, b in zip(got, want):
        if a is None or b is None:
            assert a is None and b is None
        else:
            assert a.dtype == b.dtype and torch.equal(a, b)
    # gradient clipping: mean_v to norm 1, pose predictor to 0.1, everything else untouched
    torch.manual_seed(1)
    m1, m2 = _Toy(), _Toy()
    m2.load_state_dict(m1.state_dict())
    for m in (m1, m2):
        for i, p in enumerate(q for q in m.parameters() if q.requires_grad):
            p.grad = torch.full_like(p, 0.5 + i)
    t.model, t.optim = m1, SimpleNamespace(zero_grad=lambda: None)
    ra = t.collect_grad()
    rb = ref_collect(SimpleNamespace(model=m2, optim=SimpleNamespace(zero_grad=lambda: None)))
    capsys.readouterr()
    for a, b in zip(ra, rb):
        assert float(a) == pytest.approx(float(b), rel=1e-6)
    for (n, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        if p.grad is not None:
            assert torch.allclose(p.grad, q.grad, rtol=1e-6, atol=0), n


def test_symmetry_regulariser_pieces_match_reference(monkeypatch):
    """`get_symm_rots` against model/util/symmetry.py, and the one-way chamfer reduction of the symmetry loss against the
    reference's model/util/chamfer.py::chamfer_distance_single_way.  That function delegates the nearest-neighbour search
    to pytorch3d (absent offline): it is given a brute-force `knn_points` with pytorch3d's documented contract
    (K nearest points of the second cloud, SQUARED distances) -- the reduction logic around it is the reference's own.
    The surface sampling of the regulariser (pytorch3d.ops.sample_points_from_meshes, random) stays unpinned."""
    import sys
    import types
    from collections import namedtuple
    from self_corr_pose_b200.model.module import mesh as M
    ref_symm = _load('model/util/symmetry.py', 'ref_symmetry')
    for division in (1, 2, 17):
        assert torch.equal(M.get_symm_rots(division), ref_symm.get_symm_rots(division))

    KNN = namedtuple('KNN', 'dists idx knn')

    def knn_points(p1, p2, lengths1=None, lengths2=None, K=1):
        d = torch.cdist(p1.double(), p2.double()).pow(2)
        dists, idx = d.topk(K, dim=2, largest=False)
        return KNN(dists.to(p1.dtype), idx, None)

    knn = types.ModuleType('pytorch3d.ops.knn')
    knn.knn_points, knn.knn_gather = knn_points, None
    pcl = types.ModuleType('pytorch3d.structures.pointclouds')
    pcl.Pointclouds = type('Pointclouds', (), {})
    for name, mod in (('pytorch3d', types.ModuleType('pytorch3d')), ('pytorch3d.ops', types.ModuleType('pytorch3d.ops')),
                      ('pytorch3d.structures', types.ModuleType('pytorch3d.structures')), ('pytorch3d.ops.knn', knn),
                      ('pytorch3d.structures.pointclouds', pcl)):
        monkeypatch.setitem(sys.modules, name, mod)
    ref_chamfer = _load('model/util/chamfer.py', 'ref_chamfer')
    g = torch.Generator().manual_seed(2)
    x, y = torch.randn(5, 70, 3, generator=g), torch.randn(5, 400, 3, generator=g)
    want = ref_chamfer.chamfer_distance_single_way(x, y)[0]
    for chunk in (2, 16):
        got = M.chamfer_single_way(x, y, chunk=chunk)
        assert got.shape == want.shape and abs(float(got) - float(want)) < 1e-5 * float(want)


def test_camera_loss_matches_reference(monkeypatch):
    """compute_camera_loss (optional term of MeshNet.forward, flag `camera_loss`) against the reference's function,
    extracted from model/util/loss_utils.py without importing the module (it pulls in soft_renderer and pytorch3d)."""
    import ast
    from self_corr_pose_b200.model.util.loss_utils import compute_camera_loss
    tree = ast.parse(open(os.path.join(REF, 'model', 'util', 'loss_utils.py')).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'compute_camera_loss')
    ns = {'torch': torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'reference_loss_utils', 'exec'), ns)
    g = torch.Generator().manual_seed(5)
    q1, q2 = torch.linalg.qr(torch.randn(9, 3, 3, generator=g))[0], torch.linalg.qr(torch.randn(9, 3, 3, generator=g))[0]
    q2[0] = q1[0]                                           # identical rotations: clamped cosine, angle 0
    a = q1.clone().requires_grad_(True)
    b = q1.clone().requires_grad_(True)
    got, want = compute_camera_loss(a, q2), ns['compute_camera_loss'](b, q2)
    assert torch.equal(got, want)
    got[1:].sum().backward()
    want[1:].sum().backward()
    assert torch.allclose(a.grad, b.grad, rtol=1e-6, atol=1e-7)
