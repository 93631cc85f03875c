"""Legacy reference SoftRas kernel (R-GPU, baseline/_ref) vs this package's operator on a B200: CUDA-event times of the
model's four renders, forward and backward, at B x size^2 on the 1280-vertex sphere (configs[2] shape by default).
Prints a JSON object and a markdown table (copied into profiles/)."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import _rgpu, _scenes
from self_corr_pose_b200.soft_renderer import functional as srf

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mesh = sys.argv[3] if len(sys.argv) > 3 else 'uv1280'
fv, sv, f = _scenes.config0(mesh, B=B)
texs = {'mask': (torch.ones(B, fv.shape[1], 1, 3), 'surface'),
        'softtex': (srf.face_vertices(_scenes.vertex_colors(sv), f), 'vertex'),
        'depth': (srf.face_vertices(sv, f), 'vertex'),
        'hardtex': (srf.face_vertices(_scenes.vertex_colors(sv), f), 'vertex')}


def ev_time(fn, n):
    # the legacy kernels launch on the legacy default stream, which synchronises with torch's (blocking) current stream:
    # events recorded on the current stream bracket them
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
for kind, cfg in _scenes.RENDER_CONFIGS.items():
    tex, ttype = texs[kind]
    kw = dict(image_size=size, texture_type=ttype, **cfg)
    fv_d, tex_d = fv.cuda(), tex.cuda()
    g = torch.randn(B, 4, size, size, device='cuda')
    m = _rgpu.module()
    sc = _rgpu.scalars(**kw)
    col_r, info_r, aggr_r = _rgpu.forward(fv_d, tex_d, **kw)
    bufs = [torch.zeros_like(info_r), torch.zeros_like(aggr_r), torch.ones_like(col_r)]
    gbuf = [torch.zeros_like(fv_d), torch.zeros_like(tex_d)]

    def leg_f():
        m.forward_soft_rasterize(fv_d, tex_d, bufs[0], bufs[1], bufs[2], *sc)

    def leg_b():
        m.backward_soft_rasterize(fv_d, tex_d, col_r, info_r, aggr_r, gbuf[0], gbuf[1], g, *sc)
    leg_f(); leg_b()
    r = dict(legacy_fwd_ms=ev_time(leg_f, 3), legacy_bwd_ms=ev_time(leg_b, 3))
    a = fv_d.clone().requires_grad_(True)
    t = tex_d.clone().requires_grad_(True)
    out = srf.soft_rasterize(a, t, **kw)
    for _ in range(3):
        srf.soft_rasterize(a, t, **kw)
        torch.autograd.grad(out, [a, t], g, retain_graph=True, allow_unused=True)
    r['ours_fwd_ms'] = ev_time(lambda: srf.soft_rasterize(a, t, **kw), 10)
    r['ours_bwd_ms'] = ev_time(lambda: torch.autograd.grad(out, [a, t], g, retain_graph=True, allow_unused=True), 10)
    r['speedup_fwd'] = r['legacy_fwd_ms'] / r['ours_fwd_ms']
    r['speedup_bwd'] = r['legacy_bwd_ms'] / r['ours_bwd_ms']
    res[kind] = r
print(json.dumps(dict(B=B, size=size, mesh=mesh, nf=int(fv.shape[1]), renders=res)))
print('| render | legacy fwd ms | ours fwd ms | x | legacy bwd ms | ours bwd ms | x |')
print('|---|---|---|---|---|---|---|')
for k, r in res.items():
    print('| %s | %.3f | %.3f | %.1f | %.3f | %.3f | %.1f |' % (k, r['legacy_fwd_ms'], r['ours_fwd_ms'], r['speedup_fwd'],
                                                                r['legacy_bwd_ms'], r['ours_bwd_ms'], r['speedup_bwd']))
