"""One training-shape correspondence forward per path (tcgen05 / mma.sync) for an ncu capture (tools/r2_call33.sh)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_corr_pose_b200.ops.corr_match import corr_match                      # noqa: E402
from self_corr_pose_b200.model.module.correspondence import make_meshgrid     # noqa: E402

B, hf, wf, N = 64, 64, 64, 1280
g = torch.Generator().manual_seed(0)
img = F.normalize(torch.randn(B, 64, hf * wf, generator=g), 2, 1).cuda()
mesh = F.normalize(torch.relu(torch.randn(B, N, 64, generator=g)), 2, -1).cuda()
yy, xx = torch.meshgrid(torch.linspace(-1, 1, hf), torch.linspace(-1, 1, wf), indexing='ij')
mask = torch.stack([(((xx - 0.002 * b) ** 2 + yy ** 2) < 0.72 ** 2).float() for b in range(B)]).reshape(B, -1).cuda()
v = torch.randn(B, N, 3, generator=g).cuda()
grid = make_meshgrid(hf, wf, 'cuda')
with torch.no_grad():
    for mode in ('tcgen05', 'tcgen05', 'legacy', 'legacy'):
        if mode == 'legacy':
            os.environ['SCP_CORR_FWD'] = 'legacy'
        else:
            os.environ.pop('SCP_CORR_FWD', None)
        corr_match(img, mesh, mask, v, grid, 10.0, hf, wf, want_full=False, want_pool=True)
        torch.cuda.synchronize()
