"""Data-parallel plumbing: ONE flat-buffer gradient all-reduce per step (torch.distributed; NCCL over
NVLink 5 / NVSwitch on the GPUs, gloo in the CPU tests).

The reference wraps the model in DDP but then calls the inner module (model/trainer.py:70-76,121), which
disarms DDP's gradient hooks, so its ranks never average gradients.  This implements the intended
semantics: after backward, every rank holds the arithmetic mean of the per-rank gradients.

Layout: after the FIRST backward of a run the reducer adopts the gradients: every parameter that received one gets its
`.grad` re-pointed at a view of one contiguous fp32 buffer (58 MB for the full model), and autograd accumulates into
those views in place from then on -- no per-parameter pack / unpack copies around the collective, and `zero_()` is one
memset.  Parameters the graph never reaches keep `.grad is None` on every rank, exactly as with one GPU (AdamW skips
them in both cases; the graph is static, so the set is the same on all ranks and steps).
"""
import time

import torch
import torch.distributed as dist


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class FlatGradReducer:

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None            # allocated by adopt()
        self.used = []
        self.numel = 0
        self.last_ms = None         # device time of the last all-reduce (CUDA), for bench.py

    def adopt(self):
        """Moves the existing gradients into one flat buffer and re-points every .grad at its slice."""
        used = [p for p in self.params if p.grad is not None]
        self.numel = sum(p.numel() for p in used)
        dev = used[0].device if used else 'cpu'
        flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in used:
            n = p.numel()
            view = self._view(flat, p, off)
            view.copy_(p.grad)
            p.grad = view
            off += n
        self.flat, self.used = flat, used
        return self

    @staticmethod
    def _view(flat, p, off):
        """Slice [off, off + numel) of the flat buffer with the parameter's own (dense) strides: autograd's layout contract
        and the fused optimiser kernels want gradient and parameter laid out alike (channels-last convolution weights)."""
        return torch.as_strided(flat, p.shape, p.stride(), storage_offset=off)

    @property
    def adopted(self):
        return self.flat is not None

    def covers(self, model):
        """True when every gradient of `model` lives in the flat buffer (so a scan of the buffer is a scan of all of them)."""
        mine = {id(p) for p in self.used}
        return all(id(p) in mine for p in model.parameters() if p.grad is not None)

    def zero_(self):
        """Replacement of optimizer.zero_grad() once adopted: one memset, the views stay attached."""
        if self.flat is not None:
            self.flat.zero_()
            for p in self.used:         # a foreign zero_grad(set_to_none=True) may have detached a view: re-attach
                if p.grad is None:
                    self.adopt_missing()
                    break

    def adopt_missing(self):
        off = 0
        for p in self.used:
            n = p.numel()
            if p.grad is None:
                p.grad = self._view(self.flat, p, off)
            off += n

    def reduce(self, group=None, timed=False):
        """Mean over ranks of every adopted gradient: one all_reduce(SUM) of the flat buffer + one scale."""
        world = _world(group)
        if world == 1:
            return
        if self.flat is None:
            self.adopt()
        if self.numel == 0:
            return
        if timed and self.flat.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            e1.record()
            self._events = (e0, e1)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)

    def last_allreduce_ms(self):
        ev = getattr(self, '_events', None)
        if ev is None:
            return None
        ev[1].synchronize()
        return ev[0].elapsed_time(ev[1])


def shard_batch(global_batch_size, repeat, rank, world):
    """Contiguous per-rank block of whole videos: rank r owns images [r*B/world, (r+1)*B/world), keeping the
    (video, frame-slot) structure that divide_by_* relies on."""
    if global_batch_size % (world * repeat) != 0:
        raise ValueError('global batch %d is not a multiple of world*repeat = %d' % (global_batch_size, world * repeat))
    per = global_batch_size // world
    return range(rank * per, (rank + 1) * per)
