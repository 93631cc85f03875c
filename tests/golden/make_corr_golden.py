"""Generates tests/golden/corr_golden.npz by running the REFERENCE's own modules
(model/module/correspondence.py, model/module/pretrained_corr.py, model/module/network/dino.py and the zsp
ViT) on the CPU in the build container, with their unavailable dependencies stubbed out:
  * soft_renderer / pytorch3d / model.util.chamfer are imported by model/util/loss_utils.py but not used by
    the functions exercised here -> empty stub modules;
  * `.cuda()` (hard-coded in the constructors) -> identity;
  * pretrain/dino_deitsmall8_pretrain.pth -> synthetic weights from
    self_corr_pose_b200.model.module.network.vit_weights.synthetic_state_dict(seed=0).
Run:  python tests/golden/make_corr_golden.py      (needs /root/reference; the .npz is committed)
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'third-party'))

for name in ['soft_renderer', 'pytorch3d', 'pytorch3d.structures', 'pytorch3d.loss', 'pytorch3d.ops',
             'pytorch3d.ops.knn', 'pytorch3d.structures.pointclouds']:
    sys.modules[name] = types.ModuleType(name)
sys.modules['pytorch3d.ops.knn'].knn_gather = sys.modules['pytorch3d.ops.knn'].knn_points = None
sys.modules['pytorch3d.structures.pointclouds'].Pointclouds = None
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.empty_cache = lambda: None

from self_corr_pose_b200.model.module.network.vit_weights import synthetic_state_dict  # noqa: E402

tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, 'pretrain'))
torch.save(synthetic_state_dict(0), os.path.join(tmp, 'pretrain', 'dino_deitsmall8_pretrain.pth'))
os.chdir(tmp)

from model.module.correspondence import Correspondence  # noqa: E402
from model.module.pretrained_corr import PretrainedCorrespondence  # noqa: E402
import torchvision  # noqa: E402
from torchvision.transforms import InterpolationMode  # noqa: E402

out = {}
g = torch.Generator().manual_seed(7)
randn = lambda *s: torch.randn(*s, generator=g)
rand = lambda *s: torch.rand(*s, generator=g)

# ---- (1) Correspondence.match ------------------------------------------------------------------
B, hf, wf, N, C, H = 2, 16, 16, 70, 64, 64
opts = types.SimpleNamespace(tau_img=10., tau_mesh=10., topk_img=100, topk_mesh=100, corr_h=hf, corr_w=wf,
                             train=True, n_corr_feat=C, img_size=H, pretrain_k=20, divide_fn='both',
                             batch_size=2, repeat=2)
corr = Correspondence(opts)
img_feat = torch.nn.functional.normalize(randn(B, C, hf * wf), 2, 1)
mesh_feat = torch.nn.functional.normalize(torch.relu(randn(B, N, C)), 2, -1)
yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, H), indexing='ij')
mask = ((xx ** 2 + yy ** 2) < 0.6 ** 2).float()[None].repeat(B, 1, 1)
mask[1] = ((xx - 0.2) ** 2 + yy ** 2 < 0.5 ** 2).float()
pred_v = randn(B, N, 3)
pointcorr, match, imatch, conf = corr.match(img_feat, mesh_feat, mask, pred_v)
out.update(m_img_feat=img_feat, m_mesh_feat=mesh_feat, m_mask=mask, m_pred_v=pred_v,
           m_pointcorr=pointcorr.contiguous(), m_match=match, m_imatch=imatch)
assert conf is None

# ---- (2) compute_rotation_cycle_loss --------------------------------------------------------------
src_img = rand(B, 3, H, H)
tgt_feat_raw = randn(B, C, hf, wf)


class FakeEncoder:
    def encode_img(self, img):
        return None, tgt_feat_raw


torch.manual_seed(11)
loss, cycle_match, cycle_match_gt, tgt_mask_down = corr.compute_rotation_cycle_loss(src_img, mask, img_feat,
                                                                                   FakeEncoder())
torch.manual_seed(11)
angle = torch.empty(1).uniform_(0., 360.).item()
out.update(r_src_img=src_img, r_tgt_feat_raw=tgt_feat_raw, r_angle=np.float64(angle), r_loss=loss,
           r_cycle_match=cycle_match, r_cycle_match_gt=cycle_match_gt, r_tgt_mask_down=tgt_mask_down)

# ---- (3) PretrainedCorrespondence: DINO features, match, cycle loss -------------------------------
Bp = opts.batch_size * opts.repeat
pc_net = PretrainedCorrespondence(opts, mesh=None)
img = rand(Bp, 3, H, H)
maskp = mask.repeat(2, 1, 1)
depth_weight = rand(Bp, N)
img_feat_p = torch.nn.functional.normalize(randn(Bp, C, hf * wf), 2, 1)
mesh_feat_p = torch.nn.functional.normalize(torch.relu(randn(Bp, N, C)), 2, -1)
pointcorr_p, _, _, _ = corr.match(img_feat_p, mesh_feat_p, maskp, randn(Bp, N, 3))
with torch.no_grad():
    feats = pc_net.net(img)
cyc, pts_src, pts_tgt, matchp, mask_k, img_src, img_tgt = pc_net.compute_cycle_loss(img, maskp, depth_weight,
                                                                                   pointcorr_p)
out.update(p_img=img, p_mask=maskp, p_depth_weight=depth_weight, p_pointcorr=pointcorr_p.contiguous(),
           p_dino_feat=feats, p_loss=cyc, p_pts_src=pts_src, p_pts_tgt=pts_tgt, p_match=matchp, p_mask_k=mask_k)

# tiny ViT (same class, small dims) for a weights-included check of the oracle ViT restatement
from zsp.zsp.method import vision_transformer_flexible as vits  # noqa: E402
from functools import partial  # noqa: E402
torch.manual_seed(3)
tiny = vits.VisionTransformer(img_size=[32], patch_size=8, embed_dim=32, depth=3, num_heads=2, mlp_ratio=4,
                              qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6)).eval()
x = rand(2, 3, 48, 48)
with torch.no_grad():
    d = tiny.get_specific_tokens(x, layers_to_return=(1, 2))
out.update({'t_x': x, 't_k1': d[1]['k'], 't_tok2': d[2]['t']})
for k, v in tiny.state_dict().items():
    out['tw_' + k] = v

np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'corr_golden.npz'),
                    **{k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
print({k: tuple(np.shape(v)) for k, v in out.items() if not k.startswith('tw_')})
