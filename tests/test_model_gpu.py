"""GPU: the full model graph (MeshNet.forward) and the training step (Trainer.step, row a15)
on the native hot path: runs end to end, produces the reference's aux_output keys, finite losses and gradients,
parameters move, evaluation mode returns the reference's 10-tuple."""
import pytest
import torch

from self_corr_pose_b200 import synthetic
from self_corr_pose_b200.hotpath import default_opts

pytestmark = pytest.mark.gpu

KEYS = {'total_loss', 'mask_loss', 'triangle_loss', 'deform_loss', 'pullfar_loss', 'symmetry_loss', 'match_loss',
        'texture_loss', 'imatch_loss', 'cycle_loss_pretrain', 'cycle_loss', 'depth_loss'}


def test_trainer_step_and_eval_forward():
    from self_corr_pose_b200.model.trainer import Trainer
    from self_corr_pose_b200.model.module.renderer import Renderer
    torch.manual_seed(0)
    opts = default_opts(batch_size=1, repeat=4, total_iters=100)
    tr = Trainer(opts)
    model = tr.define_model()
    v, f = synthetic.load_prior('laptop')
    batch = synthetic.make_trainer_batch(opts, v, f, 4, device=tr.device, seed=0, renderer=Renderer(opts, model.mesh))
    before = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
    losses = []
    for _ in range(2):
        total, aux, grad = tr.step(batch)
        losses.append(float(total))
        assert set(aux.keys()) == KEYS
        assert all(torch.isfinite(x).all() for x in aux.values())
    moved = [n for n, p in model.named_parameters() if p.requires_grad and not torch.equal(p.detach(), before[n])]
    assert any('mean_v' in n for n in moved) and any('backbone' in n for n in moved) and any('featnet' in n for n in moved)
    assert not any('pretrain_corr_net' in n for n in moved)
    print('PARITY model step losses', losses, 'params moved', len(moved))

    opts.train = False
    model.eval()
    with torch.no_grad():
        out = model(tr.batch_reshape(batch))
    assert len(out) == 10
    pred_v, faces, tex, imatch, match, match_conf, rotation, translation, scale, pointcorr = out
    assert pointcorr.shape == (4, 4096, 995) and match.shape == (4, 3, 256, 256) and match_conf.shape == (4, 1, 256, 256)
    assert imatch.shape == (4, 2, 995) and rotation.shape == (4, 3, 3) and translation.shape == (4, 1, 3)
    det = torch.det(rotation)
    assert torch.allclose(det.abs(), torch.ones_like(det), atol=1e-4)


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize('size,corr,B,k', [(128, 32, 4, 50), (256, 64, 8, 200)])
def test_trainer_step_parity_vs_cpu_reference_formulation(size, corr, B, k, deterministic_topk):
    """Trainer.step (MeshNet.forward + backward + clip + AdamW; encoder x2, symmetry, rotation cycle included) on the GPU
    against oracle/trainer_cpu.py = the same step in the reference's formulation on the CPU, from the same weights, batch
    and random draws (global CPU generator: jitter parameters, rotation angle; the device-side surface samples of the
    symmetry loss are injected on both sides).  fp32 on both sides (TF32 off for this comparison).
    Tolerance: every aux_output scalar 1e-3 relative (observed <= 2.4e-4); updated parameters 1e-4 (observed <= 2e-5).
    Parameter gradients are reported and bounded per optimiser group at what the conditioning of the REFERENCE computation
    allows: its encoder runs BatchNorm on batch statistics of a handful of images and subtracts the vertex mean of the shape
    head, and its SoftRas alpha gradient is ill-conditioned at silhouette edges -- the same CPU code
    evaluated in fp32 and in fp64 already differs by 2e-3 (128 px) / 5e-3 (256 px) norm-wise on the backbone gradient and
    2e-3 / 3e-3 on the shape head (measured in the build container), before any change of convolution algorithm
    (cuDNN vs oneDNN).  Observed GPU vs CPU: pose head 8e-5..3e-4, mean_v 4e-4..2e-3, feature nets 3e-4..2e-3, shape head
    1e-3..3e-3, backbone 4e-3..2.5e-2.  The kernels of this package are held to 1e-3 on identical inputs elsewhere
    (tests/test_hotpath_gpu.py: gradients w.r.t. the encoder outputs 1e-6..5e-4)."""
    from oracle import hotpath_cpu as H
    from oracle import trainer_cpu as TC
    from types import SimpleNamespace
    from self_corr_pose_b200.model.trainer import Trainer
    from self_corr_pose_b200.model.module.renderer import Renderer
    from self_corr_pose_b200.model.module.mesh import sample_faces_and_weights
    from self_corr_pose_b200.model.module.network.vit_weights import synthetic_state_dict
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        opts = default_opts(img_size=size, corr_h=corr, corr_w=corr, batch_size=B // 2, repeat=2, pretrain_k=k,
                            total_iters=100, allow_tf32=False)
        torch.manual_seed(0)
        tr = Trainer(opts)
        model = tr.define_model()
        cpu = TC.CpuTrainer(opts, synthetic_state_dict(0), fma=True)
        cpu.load_state_from(model)
        v, f = synthetic.load_prior('laptop')
        mesh = SimpleNamespace(mean_v=torch.from_numpy(v), faces=torch.from_numpy(f), texture_type='vertex')
        with H.cpu_rasterizer():
            batch = synthetic.make_trainer_batch(opts, v, f, B, device='cpu', seed=2, renderer=Renderer(opts, mesh))
        # surface samples of the symmetry regulariser: drawn once on the CPU, injected on both sides
        kk = model.mesh.symm_rots.shape[0]
        gen_v = cpu.model.mesh.mean_v.detach()[None].repeat(kk * B, 1, 1)
        gen_f = cpu.model.mesh.faces[None].repeat(kk * B, 1, 1)
        torch.manual_seed(123)
        fi, w = sample_faces_and_weights(gen_v, gen_f, 10000)
        model.mesh.sampler = lambda vv, ff, n: (fi.cuda(), w.cuda())
        before = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad and 'pretrain' not in n}

        torch.manual_seed(5)
        cb = Trainer(opts)
        cb.device = torch.device('cpu')
        total_o, aux_o, grad_o = cpu.step(cb.batch_reshape(batch), symmetry_samples=(fi, w))
        torch.manual_seed(5)
        total, aux, grad = tr.step({kk_: (t.cuda() if kk_ not in ('center', 'length') else t) for kk_, t in batch.items()})
        torch.cuda.synchronize()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    rep = {kk_: (float(aux[kk_]), float(aux_o[kk_])) for kk_ in KEYS}
    print('PARITY trainer-step %dpx B%d losses %s' % (size, B, {kk_: '%.6g/%.6g' % vv for kk_, vv in rep.items()}))
    for kk_, (a, b) in rep.items():
        assert abs(a - b) <= 1e-3 * abs(b) + 1e-7, (kk_, a, b)
    # gradients (after clipping) per optimiser group, and the parameters after the AdamW step
    cpu_params = dict(cpu.model.named_parameters())
    groups = {}
    for n, p in model.named_parameters():
        if n not in cpu_params or p.grad is None:
            continue
        key = next((g for g in ('mean_v', 'pose_predictor', 'shape', 'featnet', 'backbone') if g in n), 'other')
        groups.setdefault(key, []).append((p.grad.detach().cpu().reshape(-1), cpu_params[n].grad.reshape(-1),
                                           p.detach().cpu().reshape(-1), cpu_params[n].detach().reshape(-1)))
    rels = {}
    for key, items in groups.items():
        ga, gb = torch.cat([i[0] for i in items]), torch.cat([i[1] for i in items])
        pa, pb = torch.cat([i[2] for i in items]), torch.cat([i[3] for i in items])
        rels[key] = (_rel(ga, gb), _rel(pa, pb))
    print('PARITY trainer-step %dpx B%d (grad rel, param rel) per group %s clip norms %s / %s'
          % (size, B, {kk_: '%.2e/%.2e' % vv for kk_, vv in rels.items()}, [float(x) for x in grad], [float(x) for x in grad_o]))
    bound = {'backbone': 5e-2, 'shape': 1e-2, 'featnet': 5e-3, 'mean_v': 5e-3, 'pose_predictor': 1e-3}
    for key, (gr, pr) in rels.items():
        assert gr < bound.get(key, 5e-3), (key, gr)
        assert pr < 1e-4, (key, pr)


def _snapshot(tr):
    """Values of everything a step changes, for an IN-PLACE restore (the captured graphs hold the tensors themselves)."""
    opt = tr.optim.optimizer
    return dict(model={k: v.detach().clone() for k, v in tr.model.state_dict().items()},
                state={p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in opt.state.items()},
                lrs=[g['lr'].detach().clone() if torch.is_tensor(g['lr']) else g['lr'] for g in opt.param_groups],
                sched=tr.optim.scheduler.state_dict().copy(), iters=tr.iters)


def _restore(tr, snap):
    opt = tr.optim.optimizer
    with torch.no_grad():
        for k, v in tr.model.state_dict().items():
            v.copy_(snap['model'][k])
        for p, st in snap['state'].items():
            for k, v in st.items():
                if torch.is_tensor(v):
                    opt.state[p][k].copy_(v)
                else:
                    opt.state[p][k] = v
        for g, lr in zip(opt.param_groups, snap['lrs']):
            if torch.is_tensor(g['lr']):
                g['lr'].copy_(lr)
            else:
                g['lr'] = lr
    tr.optim.scheduler.load_state_dict(dict(snap['sched']))
    tr.iters = snap['iters']


def test_graphed_step_equals_eager_step():
    """Trainer.capture / step_graphed (zero-grad + forward + backward as ONE CUDA graph, per-step host values through
    static device buffers) against the eager step from the same model / optimiser state and the same generator states:
    same losses (1e-5; SoftRas / correspondence gradient reductions are order-dependent run to run) and the same updated
    parameters; the schedule (loss weights) and the per-step random draws follow the iteration, not the capture."""
    import copy
    from self_corr_pose_b200.model.trainer import Trainer
    from self_corr_pose_b200.model.module.renderer import Renderer
    torch.manual_seed(0)
    opts = default_opts(batch_size=2, repeat=2, total_iters=50)
    tr = Trainer(opts)
    model = tr.define_model()
    v, f = synthetic.load_prior('laptop')
    batch = synthetic.make_trainer_batch(opts, v, f, 4, device=tr.device, seed=0, renderer=Renderer(opts, model.mesh))
    batch2 = synthetic.make_trainer_batch(opts, v, f, 4, device=tr.device, seed=1, renderer=Renderer(opts, model.mesh))
    tr.capture(batch, warmup=3)
    for trial, b in enumerate((batch, batch2)):          # second trial: new inputs through the static buffers, later iteration
        snap = _snapshot(tr)
        torch.manual_seed(11 + trial)
        torch.cuda.manual_seed(11 + trial)
        total_g, aux_g, _ = tr.step_graphed(b)
        aux_g = {k: float(x) for k, x in aux_g.items()}
        params_g = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
        _restore(tr, snap)
        torch.manual_seed(11 + trial)
        torch.cuda.manual_seed(11 + trial)
        total_e, aux_e, _ = tr.step(b)
        aux_e = {k: float(x) for k, x in aux_e.items()}
        print('PARITY graphed-vs-eager trial %d iters %d: %s' % (trial, tr.iters, {k: '%.6g/%.6g' % (aux_g[k], aux_e[k]) for k in aux_g}))
        for k in aux_g:
            tol = 2e-2 if k == 'symmetry_loss' else 1e-5      # the surface samples come from the device generator
            assert abs(aux_g[k] - aux_e[k]) <= tol * abs(aux_e[k]) + 1e-8, (k, aux_g[k], aux_e[k])
        worst = max(_rel(params_g[n], p.detach()) for n, p in model.named_parameters() if p.requires_grad)
        assert worst < 1e-4, worst


def test_side_streams_do_not_change_the_step():
    """MeshNet.forward issues the frozen ViT and the rotation-cycle loss (second encoder pass) on side streams; with both
    switched off the same step (same state, same generator states) gives the same losses and gradients."""
    import copy
    from self_corr_pose_b200.model.trainer import Trainer
    from self_corr_pose_b200.model.module.renderer import Renderer
    torch.manual_seed(0)
    opts = default_opts(batch_size=2, repeat=2, total_iters=50)
    tr = Trainer(opts)
    model = tr.define_model()
    v, f = synthetic.load_prior('laptop')
    batch = synthetic.make_trainer_batch(opts, v, f, 4, device=tr.device, seed=3, renderer=Renderer(opts, model.mesh))
    tr.step(batch)                                       # cuDNN autotuning settles
    snap = _snapshot(tr)
    out = {}
    for run, flag in enumerate((True, True, False)):      # two runs with side streams (run-to-run noise), one without
        _restore(tr, snap)
        model.overlap_vit = model.overlap_rotation = flag
        torch.manual_seed(21)
        torch.cuda.manual_seed(21)
        total, aux, _ = tr.step(batch)
        torch.cuda.synchronize()
        out[run] = ({k: float(x) for k, x in aux.items()},
                    {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})
    for k in out[0][0]:
        a, b = out[0][0][k], out[2][0][k]
        assert abs(a - b) <= 1e-5 * abs(b) + 1e-8, (k, a, b)
    # the (clipped) gradients as one vector.  The SoftRas / correspondence backward kernels accumulate with float reductions
    # whose order changes from run to run, and the reference's BatchNorm-on-batch-statistics encoder amplifies that noise
    # (section 2 of DESIGN.md): the yardstick is the difference between two IDENTICAL runs
    cat = lambda d: torch.cat([d[n].reshape(-1) for n in sorted(d)])
    noise = _rel(cat(out[0][1]), cat(out[1][1]))
    diff = _rel(cat(out[0][1]), cat(out[2][1]))
    print('PARITY side-streams on/off: total %.6g/%.6g, gradient rel %.2e (two identical runs: %.2e)'
          % (out[0][0]['total_loss'], out[2][0]['total_loss'], diff, noise))
    assert diff < max(5 * noise, 1e-5)
