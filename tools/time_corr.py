"""Correspondence forward on a B200: the tcgen05 path (csrc/scp_corr_tc.cu) against the mma.sync kernel (SCP_CORR_FWD=legacy)
at the training shapes (B = 64: P = 4096 x N = 1280 with the pooled outputs; the rotation cycle's 1024 x 1024), CUDA events."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_corr_pose_b200.ops.corr_match import corr_match                      # noqa: E402
from self_corr_pose_b200.model.module.correspondence import make_meshgrid     # noqa: E402


def case(B, hf, wf, N, want_pool, iters=30):
    g = torch.Generator().manual_seed(0)
    img = F.normalize(torch.randn(B, 64, hf * wf, generator=g), 2, 1).cuda()
    mesh = F.normalize(torch.relu(torch.randn(B, N, 64, generator=g)), 2, -1).cuda()
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, hf), torch.linspace(-1, 1, wf), indexing='ij')
    mask = torch.stack([(((xx - 0.002 * b) ** 2 + yy ** 2) < 0.72 ** 2).float() for b in range(B)]).reshape(B, -1).cuda()
    v = torch.randn(B, N, 3, generator=g).cuda()
    grid = make_meshgrid(hf, wf, 'cuda')
    out = {}
    for mode in ('legacy', 'tcgen05'):
        if mode == 'legacy':
            os.environ['SCP_CORR_FWD'] = 'legacy'
        else:
            os.environ.pop('SCP_CORR_FWD', None)
        with torch.no_grad():
            for _ in range(5):
                r = corr_match(img, mesh, mask, v, grid, 10.0, hf, wf, want_full=False, want_pool=want_pool)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                r = corr_match(img, mesh, mask, v, grid, 10.0, hf, wf, want_full=False, want_pool=want_pool)
            e1.record()
            torch.cuda.synchronize()
        out[mode] = (e0.elapsed_time(e1) / iters, r)
    d = max(float((a - b).abs().max()) for a, b in zip(out['legacy'][1], out['tcgen05'][1]) if a is not None)
    print('corr fwd B=%d P=%d N=%d pool=%d fg=%.2f: mma.sync %.3f ms, tcgen05 %.3f ms (max abs diff %.2e)' %
          (B, hf * wf, N, want_pool, float(mask.mean()), out['legacy'][0], out['tcgen05'][0], d), flush=True)


if __name__ == '__main__':
    case(64, 64, 64, 1280, True)
    case(64, 32, 32, 1024, False)
    case(32, 64, 64, 1280, True)
