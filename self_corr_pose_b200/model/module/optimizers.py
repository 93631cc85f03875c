"""AdamW + OneCycleLR with the reference's five name-keyed parameter groups (model/module/optimizers.py:5-84)."""
import torch


class Optimizers:
    GROUPS = ('mean_v', 'pose_predictor', ('shape_predictor', 'shape_code_predictor'), 'featnet', 'backbone')

    def __init__(self, opts, model):
        self.opts, self.model = opts, model
        self.total_steps = opts.total_iters * opts.ngpu     # sic: the reference multiplies by ngpu (appendix A.11)
        groups = [[] for _ in self.GROUPS]
        for name, p in model.named_parameters():
            if 'pretrain_corr_net' in name:
                continue
            for gi, key in enumerate(self.GROUPS):
                keys = key if isinstance(key, tuple) else (key,)
                if any(k in name for k in keys):
                    groups[gi].append(p)
                    break
        lr = opts.learning_rate
        # one multi-tensor kernel per group on the GPU (same update rule; the CPU path is pinned to the reference's optimiser).
        # capturable + one learning-rate TENSOR per group: the step can be recorded into a CUDA graph (Trainer.capture) and
        # OneCycleLR keeps writing the schedule into those tensors in place
        fused = all(p.is_cuda for g in groups for p in g) and any(len(g) for g in groups)
        if fused:
            dev = next(p.device for g in groups for p in g)
            specs = [{'params': g, 'lr': torch.tensor(lr, dtype=torch.float32, device=dev)} for g in groups]
            self.optimizer = torch.optim.AdamW(specs, lr=lr, betas=(0.9, 0.999), weight_decay=1e-4, fused=True, capturable=True)
        else:
            self.optimizer = torch.optim.AdamW([{'params': g} for g in groups], lr=lr, betas=(0.9, 0.999), weight_decay=1e-4)
        max_lrs = [opts.vert_lr_ratio * lr, opts.cam_lr_ratio * lr, lr, lr, lr]
        self.scheduler = torch.optim.lr_scheduler.OneCycleLR(self.optimizer, max_lrs, total_steps=self.total_steps,
                                                             pct_start=0.05, cycle_momentum=False, anneal_strategy='cos',
                                                             final_div_factor=25, div_factor=25)

    def step(self, iter=None):
        self.optimizer.step()
        self.scheduler.step()

    def zero_grad(self):
        self.optimizer.zero_grad()
