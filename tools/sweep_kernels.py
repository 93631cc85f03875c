"""BASELINE.json configs[4]: per-kernel sweep with CUDA events (B = 32).

SoftRas at 64/128/256/512 px x {642 V / 1280 F icosphere, 1280 V / 2556 F UV sphere, 2562 V / 5120 F icosphere} for
the soft-texture (sigma 1e-3) and depth (sigma 1e-4) renders, forward and backward; fused correspondence at
P = 256 / 1024 / 4096 patches x 1280 vertices.  Algorithmic bytes per unit are SURVEY.md section 8d's.
Writes a markdown table (argv[1], default gpurun_out/sweep.md)."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tests import _scenes
from self_corr_pose_b200 import synthetic
from self_corr_pose_b200.soft_renderer import functional as srf
from self_corr_pose_b200.ops.corr_match import corr_match
from self_corr_pose_b200.model.module.correspondence import make_meshgrid

out_path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/sweep.md'
B = 32
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    which = 'measured'
except Exception:
    HBM, which = 6650.0, 'fallback'


def timeit(fn, n=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


lines = ['# Per-kernel sweep (BASELINE configs[4]), B = %d, CUDA events, 1xB200; HBM peak %.0f GB/s (%s)\n' % (B, HBM, which),
         '## SoftRas (pack + raster kernel per launch)\n',
         '| render | px | V / F | fwd ms | fwd GB/s | fwd frac | bwd ms | bwd GB/s | bwd frac |', '|---|---:|---|---:|---:|---:|---:|---:|---:|']
meshes = {'ico642': synthetic.icosphere(3), 'uv1280': synthetic.uv_sphere(), 'ico2562': synthetic.icosphere(4)}
g = torch.Generator().manual_seed(0)
rot, trans = synthetic.random_poses(B, g)
for kind in ('softtex', 'depth'):
    for size in (64, 128, 256, 512):
        for name, (v, f) in meshes.items():
            fv, sv, ff = _scenes.screen_faces(v, f, rot, trans)
            tex = srf.face_vertices(_scenes.vertex_colors(sv), ff).cuda()
            fvd = fv.cuda().requires_grad_(True)
            kw = dict(image_size=size, texture_type='vertex', **_scenes.RENDER_CONFIGS[kind])
            t_f = timeit(lambda: srf.soft_rasterize(fvd, tex, **kw))
            o = srf.soft_rasterize(fvd, tex, **kw)
            go = torch.randn_like(o)
            t_b = timeit(lambda: torch.autograd.grad(o, fvd, go, retain_graph=True))
            nf = f.shape[0]
            bf, bb = B * (72 * nf + 24 * size * size), B * (144 * nf + 40 * size * size)
            lines.append('| %s | %d | %d / %d | %.3f | %.1f | %.4f | %.3f | %.1f | %.4f |' % (
                kind, size, v.shape[0], nf, t_f, bf / t_f / 1e6, bf / t_f / 1e6 / HBM, t_b, bb / t_b / 1e6, bb / t_b / 1e6 / HBM))
lines += ['', '## Fused correspondence (training variant: pooled pointcorr), N = 1280, C = 64\n',
          '| P | fwd ms | fwd GB/s | fwd frac | bwd ms | bwd GB/s | bwd frac |', '|---:|---:|---:|---:|---:|---:|---:|']
N, C = 1280, 64
for hf in (16, 32, 64):
    P = hf * hf
    a = F.normalize(torch.randn(B, C, P, device='cuda'), 2, 1).requires_grad_(True)
    m = F.normalize(torch.relu(torch.randn(B, N, C, device='cuda')), 2, -1).requires_grad_(True)
    md = (torch.rand(B, P, device='cuda') > 0.4).float()
    pv = torch.randn(B, N, 3, device='cuda')
    grid = make_meshgrid(hf, hf, 'cuda')
    fwd = lambda: corr_match(a, m, md, pv, grid, 10.0, hf, hf, want_full=False, want_pool=True)
    t_f = timeit(fwd)
    _, pool, mt, im, _A = fwd()
    gs = [torch.randn_like(pool), torch.randn_like(mt), torch.randn_like(im)]
    t_b = timeit(lambda: torch.autograd.grad([pool, mt, im], [a, m], gs, retain_graph=True))
    bf = B * (4 * (C * P + N * C + P + 3 * N) + 4 * (P * N // 4 + 2 * N + 3 * P))
    bb = B * 4 * (2 * C * P + 2 * N * C + P * N // 4 + 6 * P + 7 * N)
    lines.append('| %d | %.3f | %.1f | %.4f | %.3f | %.1f | %.4f |' % (P, t_f, bf / t_f / 1e6, bf / t_f / 1e6 / HBM, t_b,
                                                                   bb / t_b / 1e6, bb / t_b / 1e6 / HBM))
os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
open(out_path, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
