"""Pseudo-ground-truth 2D<->2D matches from frozen DINO features and the pre-training cycle loss.

API of the reference's model/module/pretrained_corr.py (`PretrainedCorrespondence(opts, mesh)`, `match` :48-104,
`compute_cycle_loss` :107-140).  Differences in HOW (results identical up to fp reassociation):
  * DINO features are computed once per UNIQUE image by the tcgen05 ViT (the reference feeds every image four
    times under divide_fn='both') and only up to the layer-9 key projection;
  * the (2B,1024,1024) `corr` matrix of :130-131 is never formed: only its k gathered columns are needed
    (note A):  match[:,j] = sum_n A[:,n] Pi[j,n] / (sum_n s[n] Pi[j,n] + 1e-5) with
    A = grid . Pm (2 x N) and s[n] = sum_p Pm[p,n] = [depth_weight_src[n] >= 0.5];
  * `pointcorr` may arrive already 2x2-averaged (pooled=True) from the fused correspondence kernel;
  * on CUDA the fw / bw arg-max matching of :85-89 runs as a batched tcgen05 GEMM with an arg-max epilogue on the
    token-major bf16 features of the unique images (`_argmatch_tokens`, scp_dino_argmatch) and the target-row part of
    the loss as fused kernels (ops/cycle_rows.py); the op-by-op CUDA statements are kept below (`fused=False`, feature
    maps whose size is not a multiple of 256 pixels) and serve as the parity reference in the tests.  There is no CPU
    path: the ViT raises on CPU tensors.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .network.dino import DINO
from .correspondence import make_meshgrid
from ...ops.cycle_rows import cycle_rows
from ..util.loss_utils import divide_by_frame, divide_by_instance, divide_by_both


def select_topk(score, k):
    """Indices of the k largest scores per row: `torch.topk(-distance, k)` of pretrained_corr.py:97.  The cycle distances
    live on a coarse grid, so many are EXACTLY equal (typically more than k pixels close their cycle with distance 0), and
    which of the tied pixels torch.topk returns is implementation-defined -- it differs between torch's CPU and CUDA
    kernels; the reference inherits whatever its device gives.  Parity tests replace this function (here and in
    oracle/corr.py) by a deterministic rule so that both sides select the same pixels."""
    return torch.topk(score, k=k, dim=1).indices


def decode_argmax(best):
    """scp_dino_argmatch packs (order-preserving similarity bits << 32 | 0xffffffff - column) into 64 bits (stored in an
    int64 tensor); 0 = no unmasked column -> index 0 (a constant row of the reference's masked similarity)."""
    idx = 0xffffffff - (best & 0xffffffff)
    return torch.where(best == 0, torch.zeros_like(idx), idx)


class PretrainedCorrespondence(nn.Module):

    def __init__(self, opts, mesh=None, pretrained=True, device=None):
        super().__init__()
        self.opts = opts
        self.mesh = mesh
        self.net = DINO().eval()
        for p in self.net.parameters():
            p.requires_grad = False
        self.img_size = opts.img_size
        self.feat_size = opts.img_size // 8
        self.tau_img, self.tau_mesh = opts.tau_img, opts.tau_mesh
        self.k = opts.pretrain_k
        self.hf, self.wf = opts.corr_h, opts.corr_w
        self.meshgrid = make_meshgrid(self.hf, self.wf, device if device is not None else 'cuda')
        try:
            self.divide_fn = {'frame': divide_by_frame, 'instance': divide_by_instance,
                              'both': divide_by_both}[opts.divide_fn]
        except KeyError:
            raise ValueError(opts.divide_fn)

    def _argmatch_tokens(self, tokens, src_idx, tgt_idx, src_mask_down, tgt_mask_down):
        """max_fw / max_bw of :88-89 through the native arg-max GEMM (csrc/scp_vit.cu: scp_dino_argmatch) on the
        token-major bf16 features of the UNIQUE images: the (pairs, hw, hw) similarity is never written."""
        from ... import _lib
        NP, npix = src_mask_down.shape
        dev = tokens.device
        L = _lib.lib()
        out = []
        for a_idx, w_idx, w_mask, a_mask in ((src_idx, tgt_idx, tgt_mask_down, src_mask_down),
                                             (tgt_idx, src_idx, src_mask_down, tgt_mask_down)):
            best = torch.empty(NP, npix, dtype=torch.int64, device=dev)
            w_mask = w_mask.float().contiguous()
            # token width 768 = split bf16 pairs of the fp32-class ViT mode (DINO precision 'x3'), 384 = plain bf16
            prec = _lib.VIT_X3 if tokens.shape[-1] == 768 else _lib.VIT_BF16
            with torch.cuda.device(dev):
                rc = L.scp_dino_argmatch(_lib.ptr(tokens), _lib.ptr(a_idx.contiguous()), _lib.ptr(w_idx.contiguous()),
                                         _lib.ptr(w_mask), tokens.shape[0], npix, NP, prec, _lib.ptr(best),
                                         _lib.stream_ptr(dev))
            _lib.check(rc, 'scp_dino_argmatch')
            out.append(decode_argmax(best) * (a_mask > 0))
        return out[0], out[1]     # max_fw (per source pixel), max_bw (per target pixel)

    def _match_from_feats(self, src_feat, tgt_feat, src_mask, tgt_mask, grid, argmax=None):
        bsz = src_mask.shape[0]
        fs = self.feat_size
        src_mask_down = F.interpolate(src_mask[:, None], (fs, fs), mode='nearest').reshape(bsz, -1) * 1.0
        tgt_mask_down = F.interpolate(tgt_mask[:, None], (fs, fs), mode='nearest').reshape(bsz, -1) * 1.0
        if argmax is not None:
            max_fw, max_bw = argmax(src_mask_down, tgt_mask_down)
        else:
            src_feat = src_feat.reshape(*src_feat.shape[:2], -1)
            tgt_feat = tgt_feat.reshape(*tgt_feat.shape[:2], -1)
            # arg-max of the masked similarity in both directions (pretrained_corr.py:85-89) without materialising the
            # mask product: background entries are exactly -1e5 in the reference, so (a) an unmasked row/column never
            # picks a masked partner -> a large negative bias on masked partners gives the same arg-max, folded into
            # the GEMM as two extra channels; (b) a fully masked row/column is constant -> first index (0).
            neg = -1e5
            ones = torch.ones_like(src_mask_down)[:, None]
            a = torch.cat([src_feat, ones, (neg * (src_mask_down == 0))[:, None]], dim=1)           # b, C+2, hw(src)
            b = torch.cat([tgt_feat, (neg * (tgt_mask_down == 0))[:, None], ones], dim=1)           # b, C+2, hw(tgt)
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = True      # tensor-core GEMM; the features are bf16-accurate already
            try:
                pointcorr = a.permute(0, 2, 1).bmm(b)                                              # b, hw(src), hw(tgt)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
            max_bw = pointcorr.max(1).indices * (tgt_mask_down > 0)      # per target pixel: best source pixel
            max_fw = pointcorr.max(2).indices * (src_mask_down > 0)      # per source pixel: best target pixel
        max_cy = torch.gather(max_fw, -1, max_bw)
        grid = grid.reshape(bsz, 2, -1)
        match = torch.gather(grid, -1, max_bw[:, None].expand(-1, 2, -1))
        cycle = torch.gather(grid, -1, max_cy[:, None].expand(-1, 2, -1))
        distance = (cycle - grid).norm(2, 1)
        distance = distance * (tgt_mask_down > 0) + 1e5 * (tgt_mask_down == 0)
        indices = select_topk(-distance, self.k)
        match = torch.gather(match, -1, indices[:, None].expand(-1, 2, -1))
        grid_k = torch.gather(grid, -1, indices[:, None].expand(-1, 2, -1))
        match_mask = torch.gather(tgt_mask_down, -1, indices)
        indices_match = torch.gather(max_bw, -1, indices)
        return match, grid_k, indices_match, indices, match_mask

    def match(self, src_img, tgt_img, src_mask, tgt_mask, grid):
        with torch.no_grad():
            feat = self.net(torch.cat([src_img, tgt_img], dim=0))
            bsz = src_img.shape[0]
            return self._match_from_feats(feat[:bsz], feat[bsz:], src_mask, tgt_mask, grid)

    def compute_cycle_loss(self, img, mask, depth_weight, pointcorr, pooled=False, feat=None, A=None, fused=True):
        opts = self.opts
        num_verts = pointcorr.shape[-1]
        bs, rep = opts.batch_size, opts.repeat
        h2, w2 = self.hf // 2, self.wf // 2
        B = img.shape[0]
        # pairing as index vectors (divide_by_* applied to arange): big tensors are gathered, never rolled/copied
        src_idx, tgt_idx = self.divide_fn(torch.arange(B, device=img.device), bs, rep)
        mask_src, mask_tgt = mask.index_select(0, src_idx), mask.index_select(0, tgt_idx)
        dw_src, dw_tgt = depth_weight.index_select(0, src_idx), depth_weight.index_select(0, tgt_idx)
        bsz = src_idx.shape[0]
        grid = F.interpolate(self.meshgrid.reshape(2, self.hf, self.wf)[None], (h2, w2), mode='bilinear')
        grid_flat = grid.reshape(2, -1)                                         # 2, h2*w2 (same for every pair)

        with torch.no_grad():   # DINO once per unique image, then paired
            tokens = None
            if isinstance(feat, tuple):
                feat, tokens = feat
            elif feat is None:
                npix = (self.img_size // 8) ** 2
                if img.is_cuda and npix % 256 == 0:
                    feat, tokens = self.net(img, tokens=True)
                else:
                    feat = self.net(img)
            if tokens is not None:   # native arg-max GEMM on the unique images' token-major features
                pts_src, pts_tgt, indices_src, indices_tgt, mask_k = self._match_from_feats(
                    None, None, mask_src, mask_tgt, grid.expand(bsz, -1, -1, -1),
                    argmax=lambda ms, mt: self._argmatch_tokens(tokens, src_idx, tgt_idx, ms, mt))
            else:
                pts_src, pts_tgt, indices_src, indices_tgt, mask_k = self._match_from_feats(
                    feat.index_select(0, src_idx), feat.index_select(0, tgt_idx), mask_src, mask_tgt,
                    grid.expand(bsz, -1, -1, -1))

        if not pooled:  # bilinear 1/2 with align_corners=False == exact 2x2 mean
            pointcorr = F.avg_pool2d(pointcorr.permute(0, 2, 1).reshape(B, num_verts, self.hf, self.wf), 2) \
                .reshape(B, num_verts, h2 * w2).permute(0, 2, 1)
        # per unique image: Pm = softmax over pixels (used when the image is a source) -> A = grid . Pm
        # (A may come pre-computed from the fused correspondence kernel: Correspondence.pool_A)
        if A is None:
            Pm = torch.softmax(self.tau_mesh * pointcorr, dim=1)                # B, h2*w2, N
            A = torch.matmul(grid_flat[None], Pm)                               # B, 2, N
        if fused and pointcorr.is_cuda and self.tau_img == self.tau_mesh:
            # native: gathered rows, gated softmaxes, both products, normalisation, distance (csrc/scp_cycle.cu)
            pair_loss, match = cycle_rows(pointcorr.contiguous(), A, depth_weight, src_idx, tgt_idx, indices_tgt, pts_src,
                                          mask_k, self.tau_img)
            cycle_loss = pair_loss.sum() / (bsz * indices_tgt.shape[1])
        else:
            A_src = A.index_select(0, src_idx) * (dw_src[:, None] >= 0.5)
            s_src = (dw_src >= 0.5).to(pointcorr.dtype)                         # = column sums of the gated Pm
            # target rows needed: only the k gathered pixels of every pair, straight from the per-image tensor
            # (index_select on the flattened rows: its backward is an atomic index_add, not a sorting index_put)
            flat_rows = (tgt_idx[:, None] * pointcorr.shape[1] + indices_tgt).reshape(-1)
            rows = pointcorr.reshape(-1, num_verts).index_select(0, flat_rows).reshape(bsz, -1, num_verts)  # 2B, k, N
            Pi = torch.softmax(self.tau_img * rows, dim=2) * (dw_tgt[:, None] >= 0.5)
            num = torch.matmul(A_src, Pi.permute(0, 2, 1))                      # 2B, 2, k
            den = torch.matmul(s_src[:, None], Pi.permute(0, 2, 1)) + 1e-5      # 2B, 1, k
            match = num / den
            cycle_loss = ((match - pts_src).norm(2, 1) * mask_k).mean()
        return cycle_loss, pts_src, pts_tgt, match, mask_k, img.index_select(0, src_idx), img.index_select(0, tgt_idx)
