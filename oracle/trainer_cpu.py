"""One whole training step (the reference's Trainer.train loop body, model/trainer.py:118-125, around
MeshNet.forward, model/model.py:61-152) on the CPU in the reference's own formulation -- TEST INFRASTRUCTURE / CPU
BASELINE ONLY (tests, bench.py's `cpu_baseline` and `--impl reference` legs).

Composition:
  encoder (both passes)  : the torch mirror of the reference's networks (self_corr_pose_b200/model/module/encoder.py on CPU
                           tensors; pinned to the reference's own classes by tests/test_reference_encoder_cpu.py), with
                           torchvision's ColorJitter / Normalize / rotate exactly as the reference calls them
  correspondence, SoftRas, losses, DINO ViT, pre-training cycle loss : oracle/hotpath_cpu.py (reference formulation)
  symmetry regulariser   : CanonicalMesh.compute_symmetry_loss_reference (mesh.py:53-62 op by op)
  rotation-cycle loss    : oracle/corr.py::rotation_cycle (correspondence.py:76-113)
  clip + AdamW/OneCycle  : the Trainer / Optimizers host classes (pinned to the reference's, tests/test_reference_host_cpu.py)
"""
import torch
import torch.nn.functional as F
import torchvision
from torchvision.transforms import InterpolationMode

from oracle import corr as ocorr
from oracle import hotpath_cpu as H


class CpuTrainer:
    """`model` pieces on the CPU: encoder, canonical mesh, weights schedule, optimiser; step(batch) mirrors Trainer.step."""

    def __init__(self, opts, vit_sd, use_ref=False, nthreads=0, fma=False, all_vit_blocks=False):
        from self_corr_pose_b200.model.module.encoder import Encoder
        from self_corr_pose_b200.model.module.mesh import CanonicalMesh
        from self_corr_pose_b200.model.module.optimizers import Optimizers
        from self_corr_pose_b200.model.trainer import Trainer
        self.opts, self.vit_sd = opts, vit_sd
        self.kw = dict(use_ref=use_ref, nthreads=nthreads, fma=fma, all_vit_blocks=all_vit_blocks)

        class _Net(torch.nn.Module):       # submodule names as in MeshNet: the optimiser groups are keyed by them
            def __init__(self):
                super().__init__()
                self.mesh = CanonicalMesh(opts)
                self.encoder = Encoder(opts)
        self.model = _Net()
        self.model.apply(Trainer.set_bn_eval)
        self.model.train()
        self.optim = Optimizers(opts, self.model)
        self._trainer = Trainer(opts)
        self._trainer.model, self._trainer.optim = self.model, self.optim
        self.iters = 0

    def load_state_from(self, meshnet):
        """Copies the trainable state of a product MeshNet (same names) so both sides start from the same weights."""
        sd = {k: v.detach().cpu() for k, v in meshnet.state_dict().items() if k.startswith(('mesh.', 'encoder.'))}
        self.model.load_state_dict(sd, strict=False)

    def forward(self, data, symmetry_samples=None):
        opts, m = self.opts, self.model
        img, mask, depth, occ, center, length, foc, foc_crop, pp, pp_crop, indices, gt = data
        bsz = img.shape[0]
        mean_v = m.mesh.mean_v[None].repeat(bsz, 1, 1)
        faces = m.mesh.faces[None].repeat(bsz, 1, 1)
        img_feat, mesh_feat, pred_v, rotation, translation, scale = m.encoder(img, mean_v, pp_crop, foc_crop)
        total, aux = H.forward(opts, m.mesh.mean_v, m.mesh.faces, (img, mask, depth, foc_crop, pp_crop),
                               (img_feat, mesh_feat, pred_v, rotation, translation), self.vit_sd, it=self.iters, **self.kw)
        from self_corr_pose_b200.model.module.weights import Weights
        wts = Weights(opts)
        wts.schedule(self.iters)
        if symmetry_samples is not None:
            m.mesh.sampler = lambda v, f, n: symmetry_samples
        sym = wts.symmetry_wt * m.mesh.compute_symmetry_loss_reference(pred_v, faces)
        m.mesh.sampler = None
        # rotation-cycle loss (correspondence.py:76-113): angle from the global CPU generator, torchvision rotations,
        # second encoder pass, reference formulation of the masked similarity soft-max
        hf, wf = opts.corr_h, opts.corr_w
        angle = torch.empty(1).uniform_(0., 360.).item()
        grid = ocorr.meshgrid(hf, wf).reshape(2, hf, wf)[None].repeat(bsz, 1, 1, 1)
        grid = F.interpolate(grid, (hf // 2, wf // 2), mode='bilinear')
        rotate = torchvision.transforms.functional.rotate
        src_mask = mask[:, None]
        tgt_img = rotate(img, angle, interpolation=InterpolationMode.BILINEAR)
        tgt_mask = rotate(src_mask, angle, interpolation=InterpolationMode.NEAREST)
        cycle_gt = rotate(grid, angle, interpolation=InterpolationMode.NEAREST).reshape(bsz, 2, -1)
        _, tgt_feat = m.encoder.encode_img(tgt_img)
        tgt_feat = F.normalize(tgt_feat.reshape(bsz, opts.n_corr_feat, -1), 2, 1)
        cyc, _, _ = ocorr.rotation_cycle(img_feat, tgt_feat, src_mask, tgt_mask, cycle_gt, hf, wf, opts.tau_mesh)
        extra = {'symmetry_loss': sym, 'cycle_loss': cyc * wts.cycle_loss_wt}
        total = total + sum(extra.values())
        aux.update(extra)
        aux['total_loss'] = total
        return total, aux

    def step(self, data, symmetry_samples=None):
        self.optim.zero_grad()
        total, aux = self.forward(data, symmetry_samples)
        total.mean().backward()
        grad = self._trainer.collect_grad()
        self.optim.step(self.iters)
        self.iters += 1
        return total, aux, grad
