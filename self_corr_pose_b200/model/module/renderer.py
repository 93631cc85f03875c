"""Renderer: the four SoftRas renders of one training step and the quantities derived from them.

API of the reference's model/module/renderer.py:9-73 (`Renderer(opts, mesh)`, `render_all`,
`render_mean_mesh`); renderer settings are those of renderer.py:13-27.
"""
import torch
import torch.nn.functional as F

from ... import soft_renderer as sr
from ...soft_renderer import functional as srf
from ...ops.project_faces import project_faces, FaceTopology, LOOK_AT_Z
from ..util.loss_utils import render, pinhole_cam, project_to_screen


def _soft_renderer(img_size, sigma, gamma, rgb):
    return sr.SoftRenderer(image_size=img_size, sigma_val=sigma, gamma_val=gamma, camera_mode='look_at',
                           perspective=False, aggr_func_rgb=rgb, light_mode='vertex',
                           light_intensity_ambient=1., light_intensity_directionals=0.)


class Renderer:

    def __init__(self, opts, mesh, reference_launches=False):
        """reference_launches=True reproduces the reference's launch structure (separate mask render, NOCS backward):
        used by the CPU reference-formulation harness; the product shares / skips them (identical results)."""
        self.opts = opts
        self.mesh = mesh
        self.reference_launches = reference_launches
        size = opts.img_size
        self.renderer_mask = _soft_renderer(size, 1e-4, 1e-4, 'hard')
        self.renderer_depth = _soft_renderer(size, 1e-4, 1e-4, 'softmax')
        self.renderer_softtex = _soft_renderer(size, 1e-3, 1e-2, 'softmax')
        self.renderer_hardtex = _soft_renderer(size, 1e-4, 1e-3, 'hard')
        self.renderer_depth.rasterizer.background_color = [1, 1, 1]
        self.renderer_softtex.rasterizer.background_color = [1, 1, 1]
        self._topo = None

    def _topology(self, num_verts):
        if self._topo is None or self._topo.N != num_verts or self._topo.faces.device != self.mesh.faces.device:
            self._topo = FaceTopology(self.mesh.faces, num_verts)
        return self._topo

    def _fixed_camera(self):
        """True when every renderer uses the model's camera (look_at from (0,0,-(1/tan30+1)), orthographic, scale 1),
        for which look_at is the identity rotation plus a z offset -- the case the fused geometry kernel implements."""
        for r in (self.renderer_depth, self.renderer_softtex, self.renderer_hardtex):
            tr = r.transform.transformer
            eye = getattr(tr, '_eye', None)
            if r.transform.camera_mode != 'look_at' or tr.perspective or tr.viewing_scale != 1.0 or \
                    not isinstance(eye, (list, tuple)) or list(eye) != [0, 0, -LOOK_AT_Z]:
                return False
        return True

    def render_mean_mesh(self, foc_crop, pp_crop, rotation, translation):
        bsz = rotation.shape[0]
        mean_v = self.mesh.mean_v[None].repeat(bsz, 1, 1)
        faces = self.mesh.faces[None].repeat(bsz, 1, 1)
        return render(self.renderer_depth, mean_v, faces, None, foc_crop, pp_crop, rotation, translation,
                      rotation_detach=True, translation_detach=True, render_depth=True)

    def _rasterize(self, renderer, face_vertices, face_textures):
        r = renderer.rasterizer
        return srf.soft_rasterize(face_vertices, face_textures, r.image_size, r.background_color, r.near, r.far,
                                  r.fill_back, r.eps, r.sigma_val, r.dist_func, r.dist_eps, r.gamma_val,
                                  r.aggr_func_rgb, r.aggr_func_alpha, 'vertex')

    def _visibility(self, pred_v, depth_render, foc_crop, pp_crop, rotation, translation):
        """imatch_gt = projected vertices (no y-flip), depth_weight = exp(-5 relu(z_v - z_render)) (renderer.py:64-71)."""
        imatch_gt = pred_v.detach().bmm(rotation) + translation
        imatch_depth = imatch_gt[:, :, 2].clone()
        imatch_gt = pinhole_cam(imatch_gt, pp_crop, foc_crop)[:, :, :2].permute(0, 2, 1)  # b,2,n
        imatch_depth_gt = F.grid_sample(depth_render[:, None], imatch_gt.permute(0, 2, 1)[:, None],
                                        align_corners=False)[:, 0, 0]
        depth_weight = (5 * -F.relu(imatch_depth - imatch_depth_gt)).exp().detach()
        return imatch_gt, depth_weight

    def render_all_raw(self, pred_v, faces, tex, foc_crop, pp_crop, rotation, translation):
        """The renders of render_all as the raw (B,4,H,W) SoftRas outputs, for the fused image losses
        (ops/image_losses.py): returns (r_depth, r_tex, r_nocs, imatch_gt, depth_weight) with
            mask_render = depth_mask = r_depth[:, 3], depth_render = r_depth[:, 2],
            tex_render = r_tex[:, :3], tex_mask = r_tex[:, 3], match_gt = r_nocs[:, :3], match_mask = r_nocs[:, 3].
        The three renders see the same screen-space geometry (same vertices, pose and camera; the renderers differ
        only in sigma / gamma / aggregation), so the projection, the look_at transform and the face gather are done
        ONCE here instead of once per render (loss_utils.render + sr.Mesh per call in the reference, renderer.py:39-61);
        the ambient light of intensity 1 multiplies every texture by exactly 1 and is skipped."""
        if getattr(self.mesh, 'texture_type', 'vertex') != 'vertex' or tex is None:
            raise NotImplementedError('render_all_raw: vertex textures only (every shipped config)')
        if self._fixed_camera():   # native: projection + look_at offset + both face gathers, one launch (csrc/scp_geom.cu)
            sv, fv, ft_depth = project_faces(pred_v, rotation, translation, foc_crop, pp_crop,
                                             self._topology(pred_v.shape[1]))
        else:
            sv = project_to_screen(pred_v, foc_crop, pp_crop, rotation, translation)
            fv = srf.face_vertices(self.renderer_depth.transform.transformer(sv), faces)
            ft_depth = srf.face_vertices(sv, faces)
        rd, rn = self.renderer_depth.rasterizer, self.renderer_hardtex.rasterizer
        same = all(getattr(rd, k) == getattr(rn, k) for k in ('image_size', 'near', 'far', 'fill_back', 'eps', 'sigma_val',
                                                              'dist_func', 'dist_eps', 'aggr_func_alpha'))
        if same and rd.dist_func == 'euclidean' and rd.aggr_func_alpha == 'prod' and rd.aggr_func_rgb == 'softmax' \
                and rn.aggr_func_rgb == 'hard':
            r_depth, r_nocs = srf.soft_rasterize_dual(fv, ft_depth,
                                                      srf.face_vertices(pred_v.detach(), faces), rd.image_size,
                                                      rd.background_color, rn.background_color, rd.near, rd.far,
                                                      rd.fill_back, rd.eps, rd.sigma_val, rd.dist_eps, rd.gamma_val)
        else:
            r_depth = self._rasterize(self.renderer_depth, fv, ft_depth)
            r_nocs = self._rasterize(self.renderer_hardtex, fv.detach(), srf.face_vertices(pred_v.detach(), faces))
        r_tex = self._rasterize(self.renderer_softtex, fv, srf.face_vertices(tex, faces))
        depth_render = r_depth[:, 2] if self.opts.use_depth else r_depth[:, 2].detach()
        if self._fixed_camera():   # same projection, gradient to the pose only (pred_v detached, renderer.py:64)
            sv_d = project_faces(pred_v.detach(), rotation, translation, foc_crop, pp_crop)[0]
            imatch_gt = torch.stack((sv_d[:, :, 0], -sv_d[:, :, 1]), dim=1)     # b,2,n (no y flip)
            imatch_depth_gt = F.grid_sample(depth_render[:, None], imatch_gt.permute(0, 2, 1)[:, None],
                                            align_corners=False)[:, 0, 0]
            depth_weight = (5 * -F.relu(sv_d[:, :, 2] - imatch_depth_gt)).exp().detach()
        else:
            imatch_gt, depth_weight = self._visibility(pred_v, depth_render, foc_crop, pp_crop, rotation, translation)
        return r_depth, r_tex, r_nocs, imatch_gt, depth_weight

    def render_all(self, pred_v, faces, tex, foc_crop, pp_crop, rotation, translation, scale=None):
        texture_type = getattr(self.mesh, 'texture_type', 'vertex')
        # The mask, depth and NOCS renders share sigma = 1e-4 and 'prod' alpha aggregation, and the alpha channel is
        # accumulated before any RGB logic (soft_rasterize_cuda_kernel.cu:408-417): their alpha channels are the same
        # tensor (note B).  One traversal (the depth render) therefore provides mask_render too;
        # autograd sums the mask-loss and depth-loss gradients into a single backward launch.
        depth_full = render(self.renderer_depth, pred_v, faces, None, foc_crop, pp_crop, rotation, translation,
                            render_depth=True, texture_type='vertex')
        if self.reference_launches:
            mask_render = render(self.renderer_mask, pred_v, faces, None, foc_crop, pp_crop, rotation, translation,
                                 render_mask=True, texture_type='vertex')[:, -1]
        else:
            mask_render = depth_full[:, 3]
        depth_mask = depth_full[:, 3]
        depth_render = depth_full[:, 2].clone()
        if not self.opts.use_depth:
            depth_render = depth_render.detach()

        if tex is not None:
            tex_render = render(self.renderer_softtex, pred_v, faces, tex, foc_crop, pp_crop, rotation,
                                translation, texture_type=texture_type)
            tex_mask, tex_render = tex_render[:, -1], tex_render[:, :3]
        else:
            tex_mask = tex_render = None

        # NOCS map: hard RGB with the canonical (detached) coordinates as vertex colours.  Its gradient w.r.t.
        # geometry is exactly zero (hard RGB has no xyz gradient, soft_rasterize_cuda_kernel.cu:602; the alpha
        # channel only feeds a boolean mask, loss_utils.py:318), so the pose is detached too and the backward
        # launch the reference performs for this render is skipped.
        detach_pose = not self.reference_launches
        match_gt = render(self.renderer_hardtex, pred_v.detach(), faces, pred_v.detach(), foc_crop, pp_crop,
                          rotation, translation, rotation_detach=detach_pose, translation_detach=detach_pose,
                          texture_type='vertex')
        match_mask, match_gt = match_gt[:, -1], match_gt[:, :3]

        # projected vertices (no y-flip), visibility weight from the rendered depth
        imatch_gt, depth_weight = self._visibility(pred_v, depth_render, foc_crop, pp_crop, rotation, translation)
        return (mask_render, tex_render, depth_render, match_gt, imatch_gt, tex_mask, depth_mask, match_mask,
                depth_weight)
