"""One launch of each hot kernel between cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full`."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from tests import _scenes
from self_corr_pose_b200.soft_renderer import functional as srf
from self_corr_pose_b200.ops.corr_match import corr_match
from self_corr_pose_b200.model.module.correspondence import make_meshgrid
from self_corr_pose_b200.model.module.network.dino import DINO
from self_corr_pose_b200 import _lib

B = 64
fv, sv, f = _scenes.config0('uv1280', B=B)
tex = srf.face_vertices(_scenes.vertex_colors(sv), f).cuda()
fvd = fv.cuda().requires_grad_(True)
kw = dict(image_size=256, texture_type='vertex', **_scenes.RENDER_CONFIGS['softtex'])
kwd = dict(image_size=256, texture_type='vertex', **_scenes.RENDER_CONFIGS['depth'])
g = torch.randn(B, 4, 256, 256, device='cuda')
hf = wf = 64; N = 1280
img_feat = F.normalize(torch.randn(B, 64, hf * wf, device='cuda'), 2, 1).requires_grad_(True)
mesh_feat = F.normalize(torch.relu(torch.randn(B, N, 64, device='cuda')), 2, -1).requires_grad_(True)
mask_down = (torch.rand(B, hf * wf, device='cuda') > 0.4).float()
pred_v = torch.randn(B, N, 3, device='cuda')
grid = make_meshgrid(hf, wf, 'cuda')
net = DINO().cuda()
img = torch.rand(B, 3, 256, 256, device='cuda')


def run_all():
    o = srf.soft_rasterize(fvd, tex, **kw); o.backward(g)
    o = srf.soft_rasterize(fvd, tex, **kwd); o.backward(g)
    _, pp, m, im, A = corr_match(img_feat, mesh_feat, mask_down, pred_v, grid, 10.0, hf, wf, want_full=False, want_pool=True)
    torch.autograd.backward([pp, m, im, A], [torch.ones_like(pp), torch.ones_like(m), torch.ones_like(im), torch.ones_like(A)])
    net(img)


run_all(); run_all()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run_all()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
