"""Inference pose fit: 9-DoF pose of every image from the dense 2D-3D correspondences and the depth map.

SURVEY.md section 8f-2 -- behaviour of `Tester.pose_fitting` of the reference (model/tester.py:324-427) together with
the constants its `Tester.test` prepares (:131-137 pixel-centre grid, :150 base rotation).  The reference loops over the
images; per image it compacts the confident foreground pixels with boolean indexing (three host synchronisations),
back-projects them with the crop intrinsics and runs the 100-round RANSAC of model/util/umeyama.py in Python.  Here
the whole batch is compacted by one stable sort, back-projected at once and fitted by `fit_similarity_batch` (all
rounds of all images in one batched SVD + one residual table); two host synchronisations per batch (pixel counts,
RANSAC verdicts).  Same results and the same consumption of the global CPU random generator as the reference loop.

Deviation: when the reference's fit returns no transform (fewer than 10 % inliers, umeyama.py:29-31) its caller fails on
`None.reshape` outside the try block (tester.py:376); here such an image gets the same default pose as an image whose
fit raised (:370-374).
"""
import itertools

import torch

from .util.umeyama import fit_similarity_batch

# pose of an image whose fit failed (tester.py:371-373), in the millimetre units of the fit
_DEFAULT_SCALE, _DEFAULT_TRANSLATION = 100., (0., 0., 500.)
_UNIT = 0.001       # millimetres -> metres (:386-387)


class PoseFitter:

    def __init__(self, opts, device=None, base_rot=None):
        self.opts = opts
        size = opts.img_size
        dev = torch.device('cuda' if device is None else device)
        centre = (torch.arange(size, device=dev, dtype=torch.float32) + 0.5) / (size / 2) - 1
        # (h*w, 2) pixel-centre NDC coordinates, x fastest -- the reference keeps the same numbers flattened as (2*h*w,)
        self.grid = torch.stack((centre[None].expand(size, size), centre[:, None].expand(size, size)), -1).reshape(-1, 2)
        if base_rot is None:    # model/util/base_rot.py:10-18: flag `base_rot`, row-major 3x3 (identity by default)
            base_rot = [float(x) for x in getattr(opts, 'base_rot', (1, 0, 0, 0, 1, 0, 0, 0, 1))]
        self.base_rot = torch.as_tensor(base_rot, dtype=torch.float32, device=dev).reshape(1, 3, 3)

    def correspondences(self, mask, depth, match, match_conf, foc_crop, pp_crop):
        """Confident foreground pixels of every image, compacted to the front in pixel order.
        Returns (model points (B,n,3) from `match`, camera points (B,n,3) back-projected from `depth`, counts)."""
        B = mask.shape[0]
        keep = ((depth > 0)[:, None] * mask[:, None] * match_conf).reshape(B, -1) > 0
        counts = keep.sum(1).tolist()                                        # host synchronisation 1 of 2
        n = max(max(counts), 1)
        order = torch.sort(keep.to(torch.uint8), dim=1, descending=True, stable=True).indices[:, :n]
        K = torch.eye(3, device=mask.device)[None].repeat(B, 1, 1)
        K[:, 0, 0], K[:, 1, 1] = foc_crop[:, 0], foc_crop[:, 1]
        K[:, 0, 2], K[:, 1, 2] = pp_crop[:, 0], pp_crop[:, 1]
        K_inv = K.inverse()
        px = self.grid[order]                                                # B, n, 2
        rays = torch.cat((px, torch.ones_like(px[..., :1])), -1).bmm(K_inv.transpose(1, 2))
        d = depth.reshape(B, -1).gather(1, order)[..., None]
        cam = rays * d / rays[..., 2:]
        model = match.reshape(B, 3, -1).gather(2, order[:, None].expand(-1, 3, -1)).transpose(1, 2)
        return model, cam, counts

    def fit(self, model, cam, counts):
        """(rotation (B,3,3), translation (B,1,3), scale (B,1,3)) in metres; default pose where the fit failed."""
        s, R, t, ok = fit_similarity_batch(model, cam, counts)
        ok = torch.tensor(ok, device=model.device)
        eye = torch.eye(3, device=model.device, dtype=model.dtype)
        R = torch.where(ok[:, None, None], R, eye)
        t = torch.where(ok[:, None], t, torch.tensor(_DEFAULT_TRANSLATION, device=model.device, dtype=model.dtype))
        s = torch.where(ok, s, torch.full_like(s, _DEFAULT_SCALE))
        return R, (t * _UNIT)[:, None], (s * _UNIT)[:, None, None].expand(-1, 1, 3)

    def pose_fitting(self, batch, pred):
        """batch: the 12-tuple of `batch_reshape`; pred: MeshNet's evaluation output.  Returns (bbox (B,9,3): centre + 8
        corners of the axis-aligned box of the aligned shape, posed; posed vertices (B,N,3); rotation; translation)."""
        img, mask, depth, occ, center, length, foc, foc_crop, pp, pp_crop, indices, *_ = batch
        pred_v, faces, tex, imatch, match, match_conf, *_ = pred
        model, cam, counts = self.correspondences(mask, depth, match, match_conf, foc_crop, pp_crop)
        rotation, translation, scale_fit = self.fit(model, cam, counts)

        B = pred_v.shape[0]
        base = self.base_rot.to(pred_v.device).expand(B, -1, -1)
        pred_v = pred_v.bmm(base.transpose(1, 2))           # shape in the axes of the ground-truth prior (:396-397)
        rotation = base.bmm(rotation)
        lo, hi = pred_v.min(1).values, pred_v.max(1).values                  # B, 3
        ends = torch.stack((lo, hi), 1)                                      # B, 2, 3
        corners = [torch.stack([ends[:, c[k], k] for k in range(3)], -1) for c in itertools.product((0, 1), repeat=3)]
        bbox = torch.stack([(lo + hi) / 2] + corners, dim=-2)                # centre, then corners with z fastest
        pose = lambda x: (x * scale_fit).bmm(rotation) + translation
        return pose(bbox), pose(pred_v), rotation, translation
