"""Dense 2D<->3D correspondence between image features and mesh-vertex features.

API of the reference's model/module/correspondence.py (`Correspondence(opts)`, `match` :36-73,
`compute_rotation_cycle_loss` :76-113); the P x N similarity, both softmaxes and the two weighted sums
run in the fused sm_100a kernels of ops/corr_match.py instead of materialised torch ops.
"""
import numpy as np
import torch
import torch.nn.functional as F
import torchvision
from torchvision.transforms import InterpolationMode

from ... import _flags
from ...ops.corr_match import corr_match


def make_meshgrid(hf, wf, device):
    """(2, hf*wf) pixel-centre grid, x fastest, BOTH axes divided by wf/2 as in correspondence.py:31-33."""
    ys, xs = torch.meshgrid(torch.arange(hf, dtype=torch.float32), torch.arange(wf, dtype=torch.float32),
                            indexing='ij')
    grid = torch.stack([xs.reshape(-1), ys.reshape(-1)], 0) + 0.5
    return (grid / (wf / 2) - 1).to(device)


class RotationSlot:
    """The rotation of the rotation-cycle loss with its affine matrix in device memory: what
    torchvision.transforms.functional.rotate(img, angle) computes for tensors (functional.py: inverse affine matrix of
    -angle about the centre; _functional_tensor.rotate: sampling grid from the matrix, grid_sample with zero padding),
    split into a host part (`refresh`: draws the angle from the global CPU generator like correspondence.py:82 and writes
    the six matrix entries into pinned memory) and a device part that reads the matrix from a static tensor, so that the
    call site can be captured into a CUDA graph and replayed with a new angle."""

    def __init__(self, device):
        self.host = _flags.pinned(torch.zeros(6, dtype=torch.float32))
        self.theta = torch.zeros(1, 2, 3, dtype=torch.float32, device=device)
        self.angle = 0.
        self._cache = {}

    def refresh(self):
        from torchvision.transforms.functional import _get_inverse_affine_matrix
        self.angle = torch.empty(1).uniform_(0., 360.).item()
        m = _get_inverse_affine_matrix([0.0, 0.0], -self.angle, [0.0, 0.0], 1.0, [0.0, 0.0])
        self.host.copy_(torch.tensor(m, dtype=torch.float32))

    def upload(self):
        self.theta.view(-1).copy_(self.host, non_blocking=True)

    def _constants(self, w, h):
        """Per image size: torchvision's base grid of pixel centres and the (w/2, h/2) divisor (_gen_affine_grid), built
        once OUTSIDE any graph capture (torchvision creates them with host -> device copies on every call)."""
        key = (w, h)
        if key not in self._cache:
            dev, d = self.theta.device, 0.5
            base = torch.empty(1, h, w, 3, dtype=torch.float32, device=dev)
            base[..., 0].copy_(torch.linspace(-w * 0.5 + d, w * 0.5 + d - 1, steps=w, device=dev))
            base[..., 1].copy_(torch.linspace(-h * 0.5 + d, h * 0.5 + d - 1, steps=h, device=dev).unsqueeze_(-1))
            base[..., 2].fill_(1)
            self._cache[key] = (base.view(1, h * w, 3), torch.tensor([0.5 * w, 0.5 * h], dtype=torch.float32, device=dev))
        return self._cache[key]

    def rotate(self, img, mode):
        from torchvision.transforms import _functional_tensor as FT
        w, h = img.shape[-1], img.shape[-2]
        base, half = self._constants(w, h)
        grid = base.bmm(self.theta.transpose(1, 2) / half).view(1, h, w, 2)       # _gen_affine_grid's statements
        return FT._apply_grid_transform(img, grid, mode, fill=None)


class Correspondence:

    def __init__(self, opts, device=None):
        self.opts = opts
        self.tau_img, self.tau_mesh = opts.tau_img, opts.tau_mesh
        if self.tau_img != self.tau_mesh:
            raise NotImplementedError('the fused correspondence kernel shares one exponential between the two '
                                      'softmaxes and needs tau_img == tau_mesh (true for every shipped config)')
        self.k_img, self.k_mesh = opts.topk_img, opts.topk_mesh
        self.hf, self.wf = opts.corr_h, opts.corr_w
        self.meshgrid = make_meshgrid(self.hf, self.wf, device if device is not None else 'cuda')
        self.pool_A = None

    def match_lowres(self, img_feat, mesh_feat, mask, pred_v):
        """Training-step variant of `match` for the fused image losses: returns (pooled pointcorr (B,P/4,N),
        match at the correspondence resolution (B,P,3) -- the loss kernel applies the nearest upsampling of :71 on
        the fly --, imatch (B,2,N)); sets `pool_A` like `match(pooled=True)`."""
        bsz = mask.shape[0]
        mask_down = F.interpolate(mask[:, None], (self.hf, self.wf), mode='nearest').reshape(bsz, -1) * 1.0
        _, pc_pool, match, imatch, A_pool = corr_match(img_feat, mesh_feat, mask_down, pred_v.detach(), self.meshgrid,
                                                       self.tau_img, self.hf, self.wf, want_full=False, want_pool=True)
        self.pool_A = A_pool
        return pc_pool, match, imatch

    def match(self, img_feat, mesh_feat, mask, pred_v, pooled=False):
        """Returns (pointcorr, match, imatch, match_conf).  `pooled=True` (used by MeshNet.forward in
        training, where the only consumer is the pre-training cycle loss) returns the 2x2-averaged
        pointcorr (B, P/4, N) instead of the full (B, P, N) map."""
        bsz, h, w = mask.shape
        opts = self.opts
        mask_down = F.interpolate(mask[:, None], (self.hf, self.wf), mode='nearest').reshape(bsz, -1) * 1.0
        pc_full, pc_pool, match, imatch, A_pool = corr_match(img_feat, mesh_feat, mask_down, pred_v.detach(),
                                                             self.meshgrid, self.tau_img, self.hf, self.wf,
                                                             want_full=not pooled, want_pool=pooled)
        pointcorr = pc_pool if pooled else pc_full
        # by-product of the pooled kernel: pooled_grid . softmax(tau * pointcorr_pool, dim=pixels) per image, the
        # source-side factor of the pre-training cycle loss (PretrainedCorrespondence.compute_cycle_loss(A=...))
        self.pool_A = A_pool

        match_conf = None if opts.train else self.match_confidence(match, imatch, pred_v, mask)

        match = F.interpolate(match.reshape(bsz, self.hf, self.wf, 3).permute(0, 3, 1, 2), (h, w), mode='nearest')
        return pointcorr, match, imatch, match_conf

    def match_confidence(self, match, imatch, pred_v, mask):
        """Forward-backward consistency confidence of the dense matches, evaluation only (correspondence.py:57-69).
        match (B,P,3): soft 3D match of every correspondence-map pixel; imatch (B,2,N): soft 2D match of every vertex.
        A pixel is confident when the 2D match of the vertex nearest to its 3D match falls back onto the pixel:
        conf = exp(-5 |pixel - imatch[nearest vertex]|), upsampled bilinearly to the mask's size; values under
        min(mean over the foreground, 0.5) are zeroed.  Plain torch; the nearest-vertex search accumulates a (B,N,P)
        table of squared distances instead of the reference's (B,N,P,3) difference tensor."""
        bsz, h, w = mask.shape
        with torch.no_grad():
            d2 = None                                       # squared distances vertex <-> 3D match, one coordinate at a time
            for k in range(3):
                diff = pred_v[:, :, None, k] - match[:, None, :, k]                                       # B, N, P
                d2 = diff * diff if d2 is None else d2 + diff * diff
            near = d2.argmin(1)                                                                           # B, P
            back = torch.gather(imatch.permute(0, 2, 1), 1, near[:, :, None].expand(-1, -1, 2))           # B, P, 2
            fberr = (self.meshgrid.permute(1, 0)[None] - back).norm(2, -1).view(bsz, 1, self.hf, self.wf)
            conf = (-5 * fberr).exp()
        conf = F.interpolate(conf, (h, w), mode='bilinear', align_corners=False).detach()
        floor = min(conf[mask[:, None] > 0].mean().item(), 0.5)
        conf[conf < floor] = 0
        return conf

    def compute_rotation_cycle_loss(self, src_img, src_mask, src_img_feat, encoder):
        bsz = src_img.shape[0]
        hf2, wf2 = self.hf // 2, self.wf // 2
        grid = self.meshgrid.reshape(2, self.hf, self.wf)[None].repeat(bsz, 1, 1, 1)
        grid = F.interpolate(grid, (hf2, wf2), mode='bilinear')

        src_mask = src_mask[:, None]
        slot = getattr(self, 'rotation_slot', None)
        if slot is None:
            angle = torch.empty(1).uniform_(0., 360.).item()
            rotate = torchvision.transforms.functional.rotate
            tgt_img = rotate(src_img, angle, interpolation=InterpolationMode.BILINEAR)
            tgt_mask = rotate(src_mask, angle, interpolation=InterpolationMode.NEAREST)
            cycle_match_gt = rotate(grid, angle, interpolation=InterpolationMode.NEAREST).reshape(bsz, 2, -1)
        else:       # CUDA-graph mode: same draw, the affine matrix travels through device memory
            slot.refresh()
            slot.upload()
            tgt_img = slot.rotate(src_img, 'bilinear')
            tgt_mask = slot.rotate(src_mask, 'nearest')
            cycle_match_gt = slot.rotate(grid, 'nearest').reshape(bsz, 2, -1)

        _, tgt_img_feat = encoder.encode_img(tgt_img, pass_idx=1)
        C = self.opts.n_corr_feat
        tgt_img_feat = F.normalize(tgt_img_feat.reshape(bsz, C, -1), 2, 1)

        src_mask_down = F.interpolate(src_mask, (hf2, wf2), mode='nearest').reshape(bsz, -1) * 1.0
        tgt_mask_down = F.interpolate(tgt_mask, (hf2, wf2), mode='nearest').reshape(bsz, -1) * 1.0
        tgt_f = F.interpolate(tgt_img_feat.reshape(bsz, C, self.hf, self.wf), (hf2, wf2), mode='nearest')
        src_f = F.interpolate(src_img_feat.reshape(bsz, C, self.hf, self.wf), (hf2, wf2), mode='nearest')

        # (src pixel) x (tgt pixel) similarity, softmax over src pixels, grid-weighted sum: the same fused
        # kernel as `match` with the target pixels playing the role of the vertices
        grid_flat = grid[0].reshape(2, -1).contiguous()
        tgt_rows = tgt_f.reshape(bsz, C, -1).permute(0, 2, 1).contiguous()
        dummy_v = torch.zeros(bsz, tgt_rows.shape[1], 3, device=src_img.device)
        _, _, _, cycle_match, _ = corr_match(src_f.reshape(bsz, C, -1), tgt_rows, src_mask_down, dummy_v, grid_flat,
                                             self.tau_mesh, hf2, wf2, want_full=False, want_pool=False)
        # masked target columns: the reference's softmax is uniform there -> mean of the grid
        cycle_match = torch.where(tgt_mask_down[:, None] > 0, cycle_match,
                                  grid_flat.mean(1)[None, :, None].expand_as(cycle_match))
        cycle_loss = ((cycle_match - cycle_match_gt).norm(2, 1) * tgt_mask_down).mean()
        return cycle_loss, cycle_match, cycle_match_gt, tgt_mask_down
