"""MeshNet: the full model graph on the B200-native hot path.  API of the reference's model/model.py:41-152
(`MeshNet(opts)`, `forward(data)` -> (total_loss, aux_output) in training / the 10-tuple in evaluation,
`load_network`, `.iters`, submodules `mesh, weights, encoder, corr_net, pretrain_corr_net, renderer`).
The visualisation block (:154-307, logging only) is not reproduced."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .module.weights import Weights
from .module.mesh import CanonicalMesh
from .module.encoder import Encoder
from .module.correspondence import Correspondence
from .module.pretrained_corr import PretrainedCorrespondence
from .module.renderer import Renderer
from .util import loss_utils as L
from ..ops.image_losses import image_losses
from ..ops import peer_sync_bn


class MeshNet(nn.Module):

    # kernels of this package launched by one forward + backward of the training graph: the hot path (hotpath.HotPath:
    # SoftRas 8, correspondence 12 (tcgen05 forward: 9), ViT 68, image losses 3, geometry 4, pre-training cycle rows 2, DINO arg-match 2,
    # sparse Laplacian 2) + colour jitter 2 x 2 passes + symmetry NN fwd/bwd 2 + rotation-cycle correspondence 8 + 3
    GPU_LAUNCHES = 101 + 4 + 2 + 11

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        for flag in ('flatten_loss', 'depth_loss_chamfer', 'use_occ'):      # model.py:28,30: off in every shipped config
            if getattr(opts, flag, False):
                raise NotImplementedError('option `%s` of the reference is not implemented (it is off in every shipped '
                                          'config and outside the accelerated path)' % flag)
        self.mesh = CanonicalMesh(opts)
        self.weights = Weights(opts)
        self.encoder = Encoder(opts)
        self.corr_net = Correspondence(opts)
        self.pretrain_corr_net = PretrainedCorrespondence(opts, self.mesh, pretrained=True)
        self.renderer = Renderer(opts, self.mesh)
        self.iters = 0
        self.fused_losses = True    # False: the reference's op-by-op loss statements (parity tests)
        self.overlap_vit = True     # DINO ViT on a side stream, overlapping the encoder
        self.overlap_rotation = True    # rotation-cycle loss (second encoder pass) on a second side stream
        self._side = None
        self._side2 = None
        self.triangle_loss_fn = L.LaplacianLoss(self.mesh.mean_v, self.mesh.faces, average=True)

    def forward(self, data):
        opts, wts = self.opts, self.weights
        wts.schedule(self.iters)
        img, mask, depth, occ, center, length, foc, foc_crop, pp, pp_crop, indices, gt = data
        bsz = img.shape[0]
        mean_v = self.mesh.mean_v[None].repeat(bsz, 1, 1)
        faces = self.mesh.faces[None].repeat(bsz, 1, 1)

        # the frozen DINO features depend on the images only: issued on a side stream (a parallel branch of the step's CUDA
        # graph), so that the tensor-core GEMMs of the ViT co-run with the memory-bound BatchNorm / activation / pooling
        # kernels of the encoder; joined before the pre-training cycle loss
        feat = None
        if opts.train and self.overlap_vit and img.is_cuda:
            main = torch.cuda.current_stream(img.device)
            if self._side is None:
                self._side = torch.cuda.Stream(img.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side), torch.no_grad():
                want_tokens = (opts.img_size // 8) ** 2 % 256 == 0
                feat = self.pretrain_corr_net.net(img, tokens=want_tokens)
            for t in (feat if isinstance(feat, tuple) else (feat,)):
                t.record_stream(main)

        img_feat, mesh_feat, pred_v, rotation, translation, scale = self.encoder(img, mean_v, pp_crop, foc_crop)
        # the rotation-cycle loss (second encoder pass on the rotated images) depends on the first pass only: issued on a
        # second side stream so that its convolutions co-run with the issue-bound SoftRas / correspondence kernels of the
        # main stream, forward and -- autograd replays every node on its forward stream -- backward.  Same position in
        # the global CPU generator's sequence as in the reference (jitter, angle, jitter).
        rot_cyc = None
        if opts.train and self.overlap_rotation and img.is_cuda:
            main = torch.cuda.current_stream(img.device)
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(img.device)
            self._side2.wait_stream(main)
            with torch.cuda.stream(self._side2), peer_sync_bn.channel(1):
                rot_cyc = self.corr_net.compute_rotation_cycle_loss(img, mask, img_feat, self.encoder)
            rot_cyc[0].record_stream(main)
        fused = opts.train and self.fused_losses and opts.img_size % 16 == 0 and not opts.use_occ
        if fused:
            pointcorr, match_lr, imatch = self.corr_net.match_lowres(img_feat, mesh_feat, mask, pred_v)
        else:
            pointcorr, match, imatch, match_conf = self.corr_net.match(img_feat, mesh_feat, mask, pred_v,
                                                                       pooled=opts.train)
        tex = self.mesh.get_texture(pred_v, faces, imatch, img)
        if not opts.train:
            return pred_v, faces, tex, imatch, match, match_conf, rotation, translation, scale, pointcorr

        aux = {}
        if fused:   # shared screen-space geometry + the four image-space losses in one native forward / backward
            r_depth, r_tex, r_nocs, imatch_gt, depth_weight = self.renderer.render_all_raw(
                pred_v, faces, tex, foc_crop, pp_crop, rotation, translation)
            l_mask, l_tex, l_depth, l_match = image_losses(r_depth, r_tex, match_lr, img, mask, depth, r_nocs,
                                                           opts.corr_h, opts.corr_w, opts.use_depth)
            aux['mask_loss'] = wts.mask_wt * l_mask.mean(0)
            aux['match_loss'] = wts.match_wt * l_match.mean(0)
            aux['texture_loss'] = wts.tex_wt * l_tex.mean(0)
            if opts.use_depth:
                aux['depth_loss'] = wts.depth_wt * l_depth.mean(0)
        else:
            (mask_render, tex_render, depth_render, match_gt, imatch_gt, tex_mask, depth_mask, match_mask,
             depth_weight) = self.renderer.render_all(pred_v, faces, tex, foc_crop, pp_crop, rotation, translation, scale)
            if opts.use_occ:
                raise NotImplementedError('use_occ is False in every shipped config')
            aux['mask_loss'] = wts.mask_wt * L.compute_mask_loss(img, mask, mask_render).mean(0)
            aux['match_loss'] = wts.match_wt * L.compute_match_loss(match, match_gt, match_mask, mask).mean(0)
            aux['texture_loss'] = wts.tex_wt * L.compute_texture_loss(img, mask, tex_render, tex_mask).mean(0)
            if opts.use_depth:
                d_loss, _ = L.compute_depth_loss(depth, depth_render, depth_mask, mask)
                aux['depth_loss'] = wts.depth_wt * d_loss.mean(0)
        aux['triangle_loss'] = wts.triangle_wt * self.triangle_loss_fn(pred_v) * pred_v.shape[1] / 64.
        aux['deform_loss'] = wts.deform_wt * F.smooth_l1_loss(pred_v, mean_v, reduction='mean')
        aux['pullfar_loss'] = wts.pullfar_wt * F.relu(1 - translation[:, :, -1]).mean()
        aux['symmetry_loss'] = wts.symmetry_wt * self.mesh.compute_symmetry_loss(pred_v, faces)
        aux['imatch_loss'] = wts.imatch_wt * L.compute_imatch_loss(imatch, imatch_gt, depth_weight).mean(0)
        if feat is not None:
            torch.cuda.current_stream(img.device).wait_stream(self._side)
        cyc = self.pretrain_corr_net.compute_cycle_loss(img, mask, depth_weight, pointcorr, pooled=True,
                                                        A=self.corr_net.pool_A, feat=feat)
        aux['cycle_loss_pretrain'] = cyc[0] * wts.cycle_loss_pt_wt
        if rot_cyc is None:
            with peer_sync_bn.channel(1):
                rot_cyc = self.corr_net.compute_rotation_cycle_loss(img, mask, img_feat, self.encoder)
        else:
            torch.cuda.current_stream(img.device).wait_stream(self._side2)
        aux['cycle_loss'] = rot_cyc[0] * wts.cycle_loss_wt
        if getattr(opts, 'camera_loss', False):     # model.py:124-127: rotation against the next frame's of the same video
            nxt = rotation.detach().clone().reshape(-1, opts.repeat, 3, 3)
            nxt = torch.cat((nxt[:, 1:], nxt[:, :1]), dim=1).reshape(bsz, 3, 3)
            aux['cam_loss'] = wts.camera_wt * L.compute_camera_loss(rotation, nxt).mean()
        total_loss = sum(aux.values())
        aux_output = {'total_loss': total_loss}
        aux_output.update(aux)
        return total_loss, aux_output

    # ---- CUDA-graph support: per-step host values travel through static device buffers -------------------------------
    def enable_static_params(self, device):
        """Switches the three kinds of per-step HOST values of forward() -- the loss-weight schedule, the colour-jitter
        parameters of the two encoder passes and the angle of the rotation-cycle loss -- to static device buffers fed by
        (capturable) copies from pinned host memory.  Values and random draws are unchanged; a CUDA graph of
        forward + backward can then be replayed for new iterations: refresh_static_params(it) before every replay."""
        from .module.correspondence import RotationSlot
        self.weights.enable_device_buffer(device)
        self.encoder.enable_static_params(device)
        self.corr_net.rotation_slot = RotationSlot(device)
        self.static_params = True

    def refresh_static_params(self, it):
        """Host side of one replay: the draws of forward() in forward()'s order (jitter of the first encoder pass, rotation
        angle, jitter of the second pass: the global CPU generator is consumed exactly as by an eager step) and the
        weights of iteration `it`, written into the pinned blocks the captured copies read."""
        enc = self.encoder
        enc.jitter_slots[0].refresh(enc.random_jitter)
        self.corr_net.rotation_slot.refresh()
        enc.jitter_slots[1].refresh(enc.random_jitter)
        self.weights.fill_host(it)

    def load_network(self, model_path, iter=0):
        states = torch.load(model_path, map_location='cpu')
        for name in list(states.keys()):
            if 'symm_rots' in name or 'triangle_loss_fn' in name or 'flatten_loss_fn' in name:
                states.pop(name)
        self.load_state_dict(states, strict=False)
