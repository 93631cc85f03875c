"""Trainer: model set-up, batch reshape, one training step.  API of the reference's model/trainer.py
(`Trainer(opts)`: `define_model`, `batch_reshape`, `collect_grad`, `save`) plus `step(batch)` = the loop body
:118-125 with ONE flat-buffer gradient all-reduce between backward and clipping (the mean over ranks that the
reference's DDP wrapper intends but never performs).  TensorBoard logging / the dataset loop are
outside the hot path."""
import os

import torch

from .model import MeshNet
from .module.optimizers import Optimizers
from ..dist import FlatGradReducer


class Trainer:

    def __init__(self, opts):
        self.opts = opts
        self.iters = 0

    @staticmethod
    def set_bn_eval(m):
        if m.__class__.__name__ == 'BatchNorm2d':
            for p in m.parameters():
                p.requires_grad = False

    def define_model(self):
        opts = self.opts
        self.model = MeshNet(opts)
        if getattr(opts, 'model_path', ''):
            self.model.load_network(opts.model_path)
        self.model.apply(self.set_bn_eval)
        dev = torch.device('cuda', max(getattr(opts, 'local_rank', 0), 0))
        if getattr(opts, 'allow_tf32', True):
            # the reference's environment (torch 1.10, README.md:21) multiplies fp32 matrices and convolutions in TF32 by
            # default; torch >= 1.12 keeps that default only for cuDNN.  Restores it for the encoder's Linear layers.
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        if distributed:
            self.model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(self.model)
        self.model = self.model.to(dev)
        self.peer_bn = False
        if distributed and torch.distributed.get_world_size() > 1 and getattr(opts, 'peer_sync_bn', True):
            # statistics of the SyncBatchNorm layers over NVLink peer memory instead of 80 NCCL launches per step
            from ..ops import peer_sync_bn
            self.peer_bn = peer_sync_bn.enable(self.model, dev)
        # NHWC weights for the convolutional encoder / decoder: with the channels-last encoder input (ops/color_jitter.py)
        # cuDNN runs its tensor-core NHWC kernels end to end without per-call layout conversions (values unchanged)
        for m in (self.model.encoder.backbone, self.model.encoder.featnet):
            m.to(memory_format=torch.channels_last)
        self.device = dev
        self.optim = Optimizers(opts, self.model)
        self.reducer = FlatGradReducer([p for n, p in self.model.named_parameters() if 'pretrain_corr_net' not in n])
        self.model.train()
        return self.model

    def batch_reshape(self, batch):
        opts, dev = self.opts, self.device
        img = batch['img'].float().to(dev, non_blocking=True)
        mask = batch['mask'].to(dev, non_blocking=True).squeeze(1)
        depth = batch['depth'].to(dev, non_blocking=True).squeeze(1) if opts.use_depth else None
        occ = batch['occ'].to(dev).squeeze(1) if opts.use_occ else None
        foc, pp = batch['foc'].to(dev), batch['pp'].to(dev)
        # crop intrinsics -> NDC (fp64 end to end, trainer.py:99-101)
        pp_crop = batch['pp_crop'].to(dev) / (opts.img_size / 2.) - 1.
        foc_crop = batch['foc_crop'].to(dev) / (opts.img_size / 2.)
        return (img, mask, depth, occ, batch['center'], batch['length'], foc, foc_crop, pp, pp_crop,
                batch['idx'].to(dev), None)

    def _clip_groups(self):
        """mean_v clipped to norm 1 (its norm AFTER clipping is what the reference logs), the shape MLP to 1, the pose head
        to 0.1 (trainer.py:132-150); device ops only."""
        shapenerf, pose = [], []
        grad_meanv_norm = 0
        for name, p in self.model.named_parameters():
            if p.grad is None:
                continue
            if 'mean_v' in name:
                torch.nn.utils.clip_grad_norm_(p, 1.)
                grad_meanv_norm = p.grad.view(-1).norm(2, -1)
            elif 'shapenerf' in name:
                shapenerf.append(p)
            elif 'pose_predictor' in name:
                pose.append(p)
        return grad_meanv_norm, shapenerf, pose

    def collect_grad(self):
        grad_meanv_norm, shapenerf, pose = self._clip_groups()
        # one fused NaN check instead of a host sync per parameter (trainer.py:144-147); two launches over the flat
        # gradient buffer once the reducer has adopted the gradients, else one pair of launches per parameter
        reducer = getattr(self, 'reducer', None)
        if reducer is not None and reducer.adopted and reducer.covers(self.model):
            flat = reducer.flat.isnan().any()
        else:
            flags = [p.grad.isnan().any() for p in self.model.parameters() if p.grad is not None]
            flat = torch.stack(flags) if flags else torch.zeros(1, dtype=torch.bool)
        if bool(flat.any()):
            print('bad gradient')
            if reducer is not None and reducer.adopted:
                reducer.zero_()         # gradients live in the flat buffer: zeroed in place (views stay attached)
            else:
                self.optim.zero_grad()
        g_shape = torch.nn.utils.clip_grad_norm_(shapenerf, 1) if shapenerf else 0
        g_pose = torch.nn.utils.clip_grad_norm_(pose, 0.1) if pose else 0
        return grad_meanv_norm, g_shape, g_pose

    def step(self, batch):
        """zero_grad -> forward -> backward -> gradient all-reduce -> clip -> AdamW/OneCycle (trainer.py:118-125)."""
        self.model.iters = self.iters
        if self.reducer.adopted:        # gradients live in one flat buffer (dist.py): one memset
            self.reducer.zero_()
        else:
            self.optim.zero_grad()
        data = self.batch_reshape(batch)
        total_loss, aux_output = self.model(data)
        total_loss.mean().backward()
        self.reducer.reduce()
        grad = self.collect_grad()
        self.optim.step(self.iters)
        self.iters += 1
        return total_loss, aux_output, grad

    # ---- the step as a CUDA graph ----------------------------------------------------------------------------------
    def capture(self, batch, warmup=3):
        """Captures zero-grad + batch_reshape + forward + backward of one step into a CUDA graph over a static copy of
        `batch` (~1500 launches of this package, cuDNN and ATen per step become one graph launch; with SyncBatchNorm the
        NCCL collectives are captured too).  The gradient all-reduce, clipping (with the reference's NaN check, one host
        read) and AdamW/OneCycle stay eager after the replay.  Use step_graphed(batch) afterwards.
        Needs a few eager steps first (cuDNN autotuning, gradient buffer adoption); they are run here and DO advance the
        optimiser, like any other training step.  Call it before holding on to the loss of an eager step run on the default
        stream: a live autograd graph keeps its gradient-accumulation nodes tied to that stream, which cannot be joined to
        a capturing stream (torch raises `cudaErrorStreamCaptureImplicit`)."""
        dev = self.device
        self.model.enable_static_params(dev)
        self._static_batch = {k: (v.to(dev).clone() if torch.is_tensor(v) and k not in ('center', 'length') else v)
                              for k, v in batch.items()}
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(self._static_batch)
            if not self.reducer.adopted:
                self.reducer.adopt()        # .grad of every parameter = a view of one flat buffer: static addresses
        torch.cuda.current_stream(dev).wait_stream(side)
        self.model.iters = self.iters
        graph = torch.cuda.CUDAGraph()
        # thread_local: with a process group alive, NCCL's watchdog thread polls its events while we capture
        with torch.cuda.graph(graph, capture_error_mode='thread_local'):
            self.reducer.zero_()
            data = self.batch_reshape(self._static_batch)
            total_loss, aux_output = self.model(data)
            total_loss.mean().backward()
        self._graph = (graph, total_loss, aux_output)
        self._tail = None
        if getattr(self.opts, 'graph_optimizer', True):
            try:
                self._capture_tail()
            except Exception as e:   # noqa: BLE001 -- the eager tail of step() is always available
                print('CUDA graph capture of the clip + AdamW tail failed, keeping it eager: %r' % (e,))
                self._tail = None
        return self

    def _capture_tail(self):
        """Second graph, replayed after the gradient all-reduce: the reference's NaN rule + clipping + the fused AdamW step,
        without the host read of collect_grad.  A non-finite gradient zero-fills ALL gradients before clipping and the
        optimiser still steps -- what `self.optim.zero_grad()` of trainer.py:146 does in the reference's torch 1.10 (it
        zeroes, it does not set to None); the flag is copied to pinned memory and reported one step later.
        The graph holds the optimiser's state tensors: after `optimizer.load_state_dict` (which replaces them) capture again."""
        dev = self.device
        flat = self.reducer.flat
        self._bad_host = torch.zeros(1, dtype=torch.bool).pin_memory()
        self._bad_event = None
        for p in self.reducer.used:        # AdamW state must exist before capture (created by the warm-up steps)
            assert p in self.optim.optimizer.state, 'optimizer state missing: run eager steps before capture'
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode='thread_local'):
            bad = flat.isnan().any()
            flat.masked_fill_(bad, 0.0)
            grad_meanv_norm, shapenerf, pose = self._clip_groups()
            g_shape = torch.nn.utils.clip_grad_norm_(shapenerf, 1) if shapenerf else torch.zeros((), device=dev)
            g_pose = torch.nn.utils.clip_grad_norm_(pose, 0.1) if pose else torch.zeros((), device=dev)
            self.optim.optimizer.step()
        self._tail = (graph, bad, (grad_meanv_norm, g_shape, g_pose))

    def stage(self, batch):
        """Starts the upload of a (pinned host) batch into one of two device staging slots on a copy stream and returns
        the slot: `step_graphed(slot)` then only copies device-to-device into the graph's static buffers, so the H2D
        transfer of step i+1 overlaps the compute of step i (a data loader would call this one batch ahead)."""
        dev = self.device
        if getattr(self, '_staging', None) is None:
            keys = [k for k, v in self._static_batch.items() if torch.is_tensor(v) and k not in ('center', 'length')]
            self._staging = [{k: torch.empty_like(self._static_batch[k]) for k in keys} for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(dev)
            self._uploaded = [torch.cuda.Event() for _ in range(2)]
            self._consumed = [torch.cuda.Event() for _ in range(2)]
            for ev in self._consumed:
                ev.record(torch.cuda.current_stream(dev))
            self._stage_i = 0
        slot = self._stage_i & 1
        self._stage_i += 1
        self._copy_stream.wait_event(self._consumed[slot])      # the step that read this slot has copied it out
        with torch.cuda.stream(self._copy_stream):
            for k, dst in self._staging[slot].items():
                dst.copy_(batch[k], non_blocking=True)
            self._uploaded[slot].record(self._copy_stream)
        return slot

    def step_graphed(self, batch):
        """One training step through the captured graph: upload `batch` (a batch dict, or a slot returned by `stage`) into
        the static buffers, refresh the per-step host values, replay, then the eager tail of step()."""
        graph, total_loss, aux_output = self._graph
        if isinstance(batch, int):
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self._uploaded[batch])
            for k, v in self._staging[batch].items():
                self._static_batch[k].copy_(v, non_blocking=True)
            self._consumed[batch].record(main)
        else:
            for k, v in batch.items():
                if torch.is_tensor(v) and k not in ('center', 'length'):
                    self._static_batch[k].copy_(v, non_blocking=True)
        self.model.iters = self.iters
        self.model.refresh_static_params(self.iters)
        graph.replay()
        self.reducer.reduce()
        if self._tail is None:
            grad = self.collect_grad()
            self.optim.step(self.iters)
        else:                                   # no host synchronisation: the CPU runs ahead of the GPU
            if self._bad_event is not None and self._bad_event.query() and bool(self._bad_host):
                print('bad gradient')           # of an earlier step (its gradients were zero-filled on the device)
            tail, bad, grad = self._tail
            tail.replay()
            self._bad_host.copy_(bad.reshape(1), non_blocking=True)
            self._bad_event = torch.cuda.Event()
            self._bad_event.record(torch.cuda.current_stream(self.device))
            self.optim.scheduler.step()         # writes the next learning rates into the groups' lr tensors
        self.iters += 1
        return total_loss, aux_output, grad

    def save(self, prefix, save_dir='.'):
        if getattr(self.opts, 'local_rank', 0) <= 0:
            sd = self.model.state_dict()
            sd['mesh.faces'] = self.model.mesh.faces.cpu()
            torch.save(sd, os.path.join(save_dir, 'pred_net_{}.pth'.format(prefix)))
