"""Fused target-row part of the pre-training cycle loss (csrc/scp_cycle.cu), forward + backward.

Replaces the torch op chain of the reference's model/module/pretrained_corr.py:120-139 after the DINO matching
(row gather, gated softmaxes, the two products with the source-side factor, normalisation, distance, mean)."""
import torch
from torch.autograd import Function

from .. import _lib


def check_shapes(pc, A, dw, src_idx, tgt_idx, rows, pts_src, mask_k):
    """(B, P4, N, NP, k) after checking every buffer against the element count the kernels index."""
    B, P4, N = pc.shape
    NP, k = rows.shape
    _lib.expect_numel('cycle_rows', A_pool=(A, B * 2 * N), depth_weight=(dw, B * N), src_idx=(src_idx, NP),
                      tgt_idx=(tgt_idx, NP), pts_src=(pts_src, NP * 2 * k), mask_k=(mask_k, NP * k))
    return B, P4, N, NP, k


class CycleRowsFunction(Function):
    """(pointcorr_pool[B,P4,N], A_pool[B,2,N]; depth_weight[B,N], src_idx[NP], tgt_idx[NP], rows[NP,k] (int64),
    pts_src[NP,2,k], mask_k[NP,k], tau) -> (pair_loss[NP], match[NP,2,k])"""

    @staticmethod
    def forward(ctx, pc, A, dw, src_idx, tgt_idx, rows, pts_src, mask_k, tau):
        if not pc.is_cuda:
            raise TypeError('cycle_rows supports only CUDA tensors (no CPU path)')
        B, P4, N, NP, k = check_shapes(pc, A, dw, src_idx, tgt_idx, rows, pts_src, mask_k)
        dev = pc.device
        pc, A = pc.detach().float().contiguous(), A.detach().float().contiguous()
        dw = dw.detach().float().contiguous()
        src_idx, tgt_idx, rows = (t.detach().long().contiguous() for t in (src_idx, tgt_idx, rows))
        pts_src, mask_k = pts_src.detach().float().contiguous(), mask_k.detach().float().contiguous()
        pair_loss = torch.empty(NP, dtype=torch.float32, device=dev)
        match = torch.empty(NP, 2, k, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_cycle_rows_forward(
                _lib.ptr(pc), _lib.ptr(A), _lib.ptr(dw), _lib.ptr(src_idx), _lib.ptr(tgt_idx), _lib.ptr(rows),
                _lib.ptr(pts_src), _lib.ptr(mask_k), float(tau), B, P4, N, NP, k, _lib.ptr(pair_loss), _lib.ptr(match),
                _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_cycle_rows_forward')
        ctx.save_for_backward(pc, A, dw, src_idx, tgt_idx, rows, pts_src, mask_k)
        ctx.tau = float(tau)
        ctx.mark_non_differentiable(match)
        return pair_loss, match

    @staticmethod
    def backward(ctx, g_pair, _g_match):
        pc, A, dw, src_idx, tgt_idx, rows, pts_src, mask_k = ctx.saved_tensors
        B, P4, N = pc.shape
        NP, k = rows.shape
        dev = pc.device
        g_pc = torch.empty_like(pc)
        g_A = torch.empty_like(A)
        g_pair = g_pair.float().contiguous()
        with torch.cuda.device(dev):
            rc = _lib.lib().scp_cycle_rows_backward(
                _lib.ptr(pc), _lib.ptr(A), _lib.ptr(dw), _lib.ptr(src_idx), _lib.ptr(tgt_idx), _lib.ptr(rows),
                _lib.ptr(pts_src), _lib.ptr(mask_k), ctx.tau, B, P4, N, NP, k, _lib.ptr(g_pair), _lib.ptr(g_pc),
                _lib.ptr(g_A), _lib.stream_ptr(dev))
        _lib.check(rc, 'scp_cycle_rows_backward')
        return g_pc, g_A, None, None, None, None, None, None, None


def cycle_rows(pc, A, dw, src_idx, tgt_idx, rows, pts_src, mask_k, tau):
    return CycleRowsFunction.apply(pc, A, dw, src_idx, tgt_idx, rows, pts_src, mask_k, tau)
