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
