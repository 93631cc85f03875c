"""SyncBatchNorm whose statistics exchange runs over NVLink peer memory (csrc/scp_peer.cu) instead of NCCL.

The reference converts every BatchNorm2d of the model to torch.nn.SyncBatchNorm when it is launched distributed
(model/trainer.py:66): forward all-gathers (mean, invstd, count) of every layer, backward all-reduces (sum_dy, sum_dy_xmu).
PeerSyncBatchNorm keeps exactly those statements (torch.batch_norm_stats / batch_norm_gather_stats_with_counts /
batch_norm_elemt / batch_norm_backward_reduce / batch_norm_backward_elemt, as torch.nn.modules._functions.SyncBatchNorm)
and swaps the two collectives for `scp_peer_exchange`: one kernel that stores the vector into every peer's buffer, raises a
flag and waits for the peers' flags.  One node only (cudaIpc); on any set-up failure the model keeps torch's NCCL layers.

Channels: a channel is one peer buffer + one sequence counter; all ranks must issue the same sequence of exchanges on a
channel in stream order.  The two encoder passes of a step run on different CUDA streams (MeshNet.forward), so each pass
has its own channel (`with channel(1): ...` around the second pass); a layer's backward uses the channel of its forward.
"""
import contextlib
import ctypes

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.autograd import Function

from .. import _lib

_CHANNELS = []          # PeerChannel objects, created by enable() before any CUDA-graph capture
_CURRENT = 0


class PeerChannel:

    def __init__(self, group, device):
        L = _lib.lib()
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise RuntimeError('peer exchange supports at most 16 ranks')
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            _lib.check(L.scp_peer_buffer_create(ctypes.byref(own), handle), 'scp_peer_buffer_create')
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            ptrs = []
            for q, h in enumerate(handles):
                if q == self.rank:
                    ptrs.append(own.value)
                    continue
                p = ctypes.c_void_p()
                _lib.check(L.scp_peer_buffer_open(h, ctypes.byref(p)), 'scp_peer_buffer_open')
                ptrs.append(p.value)
        self.own = own.value
        self.peers = (ctypes.c_void_p * self.world)(*ptrs)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(group=group)        # every buffer is zeroed and mapped before the first exchange

    def exchange(self, src, reduce):
        """src: contiguous fp32 vector (n,) -> (world, n) gathered in rank order, or (n,) summed over ranks."""
        src = src.contiguous()
        n = src.numel()
        dst = torch.empty(n if reduce else self.world * n, dtype=torch.float32, device=src.device)
        with torch.cuda.device(src.device):
            rc = _lib.lib().scp_peer_exchange(self.peers, _lib.ptr(src), n, self.rank, self.world, _lib.ptr(self.counter),
                                              _lib.ptr(dst), 1 if reduce else 0, _lib.stream_ptr(src.device))
        _lib.check(rc, 'scp_peer_exchange')
        return dst if reduce else dst.view(self.world, n)


@contextlib.contextmanager
def channel(idx):
    """Exchanges of SyncBatchNorm layers called inside use channel `idx` (forward; the backward follows its forward)."""
    global _CURRENT
    prev, _CURRENT = _CURRENT, idx
    try:
        yield
    finally:
        _CURRENT = prev


class _PeerSyncBatchNormFn(Function):
    """torch.nn.modules._functions.SyncBatchNorm with the collectives over peer memory."""

    @staticmethod
    def forward(ctx, input, weight, bias, running_mean, running_var, eps, momentum, chan):
        if not (input.is_contiguous(memory_format=torch.channels_last) or input.is_contiguous()):
            input = input.contiguous()
        if weight is not None:
            weight = weight.contiguous()
        num_channels = input.shape[1]
        mean, invstd = torch.batch_norm_stats(input, eps)
        count = torch.full((1,), input.numel() // input.size(1), dtype=mean.dtype, device=mean.device)
        combined = torch.cat([mean, invstd, count], dim=0)                      # C, C, 1 -> (2C + 1)
        combined = chan.exchange(combined, reduce=False)                        # world x (2C + 1)
        mean_all, invstd_all, count_all = torch.split(combined, num_channels, dim=1)
        counts = count_all.view(-1)
        if running_mean is not None and counts.dtype != running_mean.dtype:
            counts = counts.to(running_mean.dtype)
        mean, invstd = torch.batch_norm_gather_stats_with_counts(input, mean_all, invstd_all, running_mean, running_var,
                                                                 momentum, eps, counts)
        ctx.save_for_backward(input, weight, mean, invstd, count_all.to(torch.int32))
        ctx.chan = chan
        return torch.batch_norm_elemt(input, weight, bias, mean, invstd, eps)

    @staticmethod
    def backward(ctx, grad_output):
        if not (grad_output.is_contiguous(memory_format=torch.channels_last) or grad_output.is_contiguous()):
            grad_output = grad_output.contiguous()
        saved_input, weight, mean, invstd, count_tensor = ctx.saved_tensors
        grad_input = None
        sum_dy, sum_dy_xmu, grad_weight, grad_bias = torch.batch_norm_backward_reduce(
            grad_output, saved_input, mean, invstd, weight, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
            ctx.needs_input_grad[2])
        if ctx.needs_input_grad[0]:
            num_channels = sum_dy.shape[0]
            combined = ctx.chan.exchange(torch.cat([sum_dy, sum_dy_xmu], dim=0), reduce=True)
            sum_dy, sum_dy_xmu = torch.split(combined, num_channels)
            if weight is not None and weight.dtype != mean.dtype:
                weight = weight.to(mean.dtype)
            grad_input = torch.batch_norm_backward_elemt(grad_output, saved_input, mean, invstd, weight, sum_dy, sum_dy_xmu,
                                                         count_tensor)
        if weight is None or not ctx.needs_input_grad[1]:
            grad_weight = None
        if weight is None or not ctx.needs_input_grad[2]:
            grad_bias = None
        return grad_input, grad_weight, grad_bias, None, None, None, None, None


class PeerSyncBatchNorm(nn.SyncBatchNorm):
    """Same module state and semantics as torch.nn.SyncBatchNorm (one-node process groups, 2*C+1 <= 1088)."""

    def forward(self, input):
        if not (self.training and input.is_cuda and _CHANNELS and 2 * self.num_features + 1 <= 1088):
            return super().forward(input)
        self._check_input_dim(input)
        self._check_non_zero_input_channels(input)
        factor = 0.0 if self.momentum is None else self.momentum
        if self.track_running_stats:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:
                factor = 1.0 / self.num_batches_tracked.item()
        running_mean = self.running_mean if self.track_running_stats else None
        running_var = self.running_var if self.track_running_stats else None
        return _PeerSyncBatchNormFn.apply(input, self.weight, self.bias, running_mean, running_var, self.eps, factor,
                                          _CHANNELS[_CURRENT])


def enable(model, device, group=None, n_channels=2):
    """Creates the channels (collective: call on every rank, outside any graph capture) and switches every
    torch.nn.SyncBatchNorm of `model` to PeerSyncBatchNorm.  Returns False (model untouched) when the peer buffers cannot be
    set up, e.g. across nodes or without peer access."""
    global _CHANNELS
    group = group if group is not None else dist.group.WORLD
    ok = torch.ones(1, device=device)
    try:
        chans = [PeerChannel(group, device) for _ in range(n_channels)]
    except Exception as e:   # noqa: BLE001 -- any rank failing disables the feature on all ranks
        print('peer-memory SyncBatchNorm unavailable (%r): keeping NCCL' % (e,))
        ok.zero_()
        chans = []
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if float(ok) < 1:
        return False
    _CHANNELS = chans
    for m in model.modules():
        if type(m) is nn.SyncBatchNorm:
            m.__class__ = PeerSyncBatchNorm
    return True
