"""Times the fused correspondence kernels (forward, backward) with CUDA events."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from self_corr_pose_b200.ops.corr_match import corr_match
from self_corr_pose_b200.model.module.correspondence import make_meshgrid

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hf = wf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
C = 64
img_feat = F.normalize(torch.randn(B, C, hf * wf, device='cuda'), 2, 1).requires_grad_(True)
mesh_feat = F.normalize(torch.relu(torch.randn(B, N, C, device='cuda')), 2, -1).requires_grad_(True)
mask_down = (torch.rand(B, hf * wf, device='cuda') > 0.4).float()
pred_v = torch.randn(B, N, 3, device='cuda')
grid = make_meshgrid(hf, wf, 'cuda')
res = {}
for full in (False, True):
    def fwd():
        return corr_match(img_feat, mesh_feat, mask_down, pred_v, grid, 10.0, hf, wf, want_full=full, want_pool=not full)
    pf, pp, m, im, _A = fwd()
    gpc = torch.randn_like(pf if full else pp); gm = torch.randn_like(m); gi = torch.randn_like(im)
    for _ in range(3):
        pf, pp, m, im, _A = fwd(); torch.autograd.backward([pf if full else pp, m, im], [gpc, gm, gi])
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    n = 10
    for _ in range(n):
        e[0].record(); pf, pp, m, im, _A = fwd(); e[1].record()
        torch.autograd.backward([pf if full else pp, m, im], [gpc, gm, gi]); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    res['full' if full else 'pooled'] = dict(fwd_ms=tf / n, bwd_ms=tb / n, fwd_us_img=1e3 * tf / n / B, bwd_us_img=1e3 * tb / n / B)
print(json.dumps(dict(B=B, P=hf * wf, N=N, **res)))
