"""Data-parallel plumbing: ONE flat-buffer gradient all-reduce per step (torch.distributed; NCCL over
NVLink 5 / NVSwitch on the GPUs, gloo in the CPU tests).

The reference wraps the model in DDP but then calls the inner module (model/trainer.py:70-76,121), which
disarms DDP's gradient hooks, so its ranks never average gradients.  This implements the intended
semantics: after backward, every rank holds the arithmetic mean of the per-rank gradients.
"""
import torch
import torch.distributed as dist


class FlatGradReducer:
    """Packs the gradients of `params` into one contiguous fp32 buffer, all-reduces it once, and scatters the
    averaged values back into the .grad tensors."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else 'cpu'
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)

    def reduce(self, group=None):
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if world == 1 or self.numel == 0:
            return
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(world)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self.flat[off:off + n].reshape(p.shape).clone()
            else:
                p.grad.copy_(self.flat[off:off + n].reshape(p.shape))
            off += n


def shard_batch(global_batch_size, repeat, rank, world):
    """Contiguous per-rank block of whole videos: rank r owns images [r*B/world, (r+1)*B/world), keeping the
    (video, frame-slot) structure that divide_by_* relies on."""
    if global_batch_size % (world * repeat) != 0:
        raise ValueError('global batch %d is not a multiple of world*repeat = %d' % (global_batch_size, world * repeat))
    per = global_batch_size // world
    return range(rank * per, (rank + 1) * per)
