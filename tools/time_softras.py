"""Times the four SoftRas renders of the model (forward + backward) with CUDA events."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import _scenes
from self_corr_pose_b200.soft_renderer import functional as srf

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mesh = sys.argv[3] if len(sys.argv) > 3 else 'uv1280'
fv, sv, f = _scenes.config0(mesh, B=B)
texs = {'mask': (torch.ones(B, fv.shape[1], 1, 3), 'surface'),
        'softtex': (srf.face_vertices(_scenes.vertex_colors(sv), f), 'vertex'),
        'depth': (srf.face_vertices(sv, f), 'vertex'),
        'hardtex': (srf.face_vertices(_scenes.vertex_colors(sv), f), 'vertex')}
res = {}
for kind, cfg in _scenes.RENDER_CONFIGS.items():
    tex, ttype = texs[kind]
    fv_d = fv.cuda().requires_grad_(True); tex_d = tex.cuda().requires_grad_(True)
    kw = dict(image_size=size, texture_type=ttype, **cfg)
    g = torch.randn(B, 4, size, size, device='cuda')
    def fwd():
        return srf.soft_rasterize(fv_d, tex_d, **kw)
    for _ in range(3):
        o = fwd(); o.backward(g)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 10
    tf = tb = 0.0
    for _ in range(n):
        e[0].record(); o = fwd(); e[1].record(); o.backward(g); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    res[kind] = dict(fwd_ms=tf / n, bwd_ms=tb / n, fwd_us_per_img=1e3 * tf / n / B, bwd_us_per_img=1e3 * tb / n / B)
print(json.dumps(dict(B=B, size=size, mesh=mesh, nf=int(fv.shape[1]), **{k: {a: round(b, 3) for a, b in v.items() if a.endswith("_ms")} for k, v in res.items()})))
