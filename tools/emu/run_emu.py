"""Functional check of SoftRas kernel variants WITHOUT a GPU: runs the host emulation (tools/emu/build_emu.py) of the
shipped csrc/scp_softras.cu -- default build and the given -D variants -- on a small scene and compares every output with
the C oracle (oracle/softras.py, strict build: g++ contracts no FMAs either, so the emulation follows the
strict oracle to ~1e-6 even where the algorithm is ill-conditioned).  The default build is the GPU-validated one: its emulation agreeing with the
oracle validates the emulator; a variant agreeing validates the variant's traversal / indexing / reductions.

    python tools/emu/run_emu.py [--big]               # default + fwd2px + facesmem; --big: 2556 faces (minutes)
Not a parity proof (libm instead of the GPU's fast-math units, one small scene): variants still go through the GPU tests."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import build_emu  # noqa: E402
from oracle import softras as osr  # noqa: E402
from self_corr_pose_b200 import _lib, synthetic  # noqa: E402
from self_corr_pose_b200.soft_renderer import functional as srf  # noqa: E402
from tests import _scenes  # noqa: E402

VARIANTS = [('', []), ('fwd2px', ['-DSCP_SOFTRAS_FWD_2PX=1']),
            ('facesmem', ['-DSCP_SOFTRAS_FACE_SMEM=1', '-DSCP_SOFTRAS_FACE_CTAS=10']),
            ('face16x2', ['-DSCP_SOFTRAS_FACE_BW=16', '-DSCP_SOFTRAS_FACE_SMEM=1', '-DSCP_SOFTRAS_FACE_CTAS=10']),
            ('facelinear', ['-DSCP_SOFTRAS_FACE_LINEAR=1'])]
_fp = ctypes.c_void_p


def load(path):
    lib = ctypes.CDLL(path)
    for name in ('scp_softras_workspace_bytes', 'scp_softras_forward', 'scp_softras_backward', 'scp_softras_forward_dual'):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = _lib._SIGNATURES[name]
    return lib


def p(a):
    return a.ctypes.data_as(_fp)


def scene(B, is_, seed, big=False):
    # 42 vertices / 80 faces, or the 1280-vertex / 2556-face sphere (several culling rounds, tile list rebuilt)
    verts, faces = synthetic.uv_sphere() if big else synthetic.icosphere(1)
    g = torch.Generator().manual_seed(seed)
    rot, trans = synthetic.random_poses(B, g)
    fv, sv, f = _scenes.screen_faces(verts, faces, rot, trans)
    tex = srf.face_vertices(_scenes.vertex_colors(sv), f)
    return np.ascontiguousarray(fv.numpy(), np.float32), np.ascontiguousarray(tex.numpy(), np.float32)


def emu_forward_backward(lib, fv, tex, g, is_, kw):
    B, nf = fv.shape[:2]
    sc = osr._scalars(B, nf, 3, is_, 1., 100., 1e-3, kw['sigma_val'], 'euclidean', 1e-4, kw['gamma_val'],
                      kw['aggr_func_rgb'], 'prod', 'vertex', True)
    ws_n = lib.scp_softras_workspace_bytes(B, nf)
    ws = np.zeros(ws_n, np.uint8)
    info = np.zeros((B, nf, 27), np.float32)
    aggr = np.zeros((B, 2, is_, is_), np.float32)
    col = np.ones((B, 4, is_, is_), np.float32)
    for k in range(3):
        col[:, k] *= kw['background_color'][k]
    rc = lib.scp_softras_forward(p(fv), p(tex), p(info), p(aggr), p(col), *sc, p(ws), ws_n, None)
    assert rc == 0, rc
    gf, gt = np.zeros((B, nf, 9), np.float32), np.zeros((B, nf, 3, 3), np.float32)
    rc = lib.scp_softras_backward(p(fv), p(tex), p(col), p(info), p(aggr), p(gf), p(gt), p(g), *sc, p(ws), ws_n, None)
    assert rc == 0, rc
    return col, aggr, gf.reshape(B, nf, 3, 3), gt


def emu_dual(lib, fv, tex, tex2, is_, sigma, gamma):
    B, nf = fv.shape[:2]
    ws_n = lib.scp_softras_workspace_bytes(B, nf)
    ws = np.zeros(ws_n, np.uint8)
    info = np.zeros((B, nf, 27), np.float32)
    aggr, aggr2 = np.zeros((B, 2, is_, is_), np.float32), np.zeros((B, 2, is_, is_), np.float32)
    col, col2 = np.ones((B, 4, is_, is_), np.float32), np.zeros((B, 4, is_, is_), np.float32)
    rc = lib.scp_softras_forward_dual(p(fv), p(tex), p(tex2), p(info), p(aggr), p(col), p(aggr2), p(col2), B, nf, is_, 1.,
                                      100., 1e-3, sigma, float(np.log(1. / 1e-4 - 1.)), gamma, 1, p(ws), ws_n, None)
    assert rc == 0, rc
    return col, col2


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


def main(big=False):
    B, is_ = (1, 48) if big else (2, 40)              # 40 px: ragged tiles (3 x 3 tiles of 16, the last ones partial)
    fv, tex = scene(B, is_, seed=3, big=big)
    g = np.random.RandomState(0).randn(B, 4, is_, is_).astype(np.float32)
    ok = True
    for name, defines in VARIANTS:
        t0 = time.time()
        lib = load(build_emu.build(name, defines))
        line = []
        for cfg in (('softtex', 'depth') if big else ('softtex', 'depth', 'mask')):
            kw = _scenes.RENDER_CONFIGS[cfg]
            okw = dict(image_size=is_, texture_type='vertex', **kw)
            col_o, info_o, aggr_o = osr.forward(fv, tex, fma=False, **okw)
            gf_o, gt_o = osr.backward(fv, tex, col_o, info_o, aggr_o, g, fma=False, **okw)
            col, aggr, gf, gt = emu_forward_backward(lib, fv, tex, g, is_, kw)
            errs = dict(col=rel(col, col_o), gf=rel(gf, gf_o), gt=rel(gt, gt_o.reshape(gt.shape)) if np.abs(gt_o).max() > 0 else 0.)
            lim = dict(col=1e-5, gf=1e-3, gt=1e-3)      # observed: 1e-7 .. 5e-5
            ok &= all(errs[k] <= lim[k] for k in errs)
            line.append('%s col %.1e gf %.1e gt %.1e' % (cfg, errs['col'], errs['gf'], errs['gt']))
        # fused depth + NOCS traversal against two separate oracle renders
        kd, kn = _scenes.RENDER_CONFIGS['depth'], _scenes.RENDER_CONFIGS['hardtex']
        tex2 = np.ascontiguousarray(tex[..., ::-1])
        col, col2 = emu_dual(lib, fv, tex, tex2, is_, kd['sigma_val'], kd['gamma_val'])
        c1 = osr.forward(fv, tex, fma=False, image_size=is_, texture_type='vertex', **kd)[0]
        c2 = osr.forward(fv, tex2, fma=False, image_size=is_, texture_type='vertex',
                         **dict(kn, gamma_val=kd['gamma_val']))[0]
        e1, e2 = rel(col, c1), float(np.mean(np.abs(col2 - c2) <= 1e-4 + 1e-3 * np.abs(c2)))
        ok &= e1 <= 1e-4 and e2 >= 0.999
        line.append('dual col %.1e nocs-within-tol %.4f' % (e1, e2))
        print('%-9s %s  (%.0f s)' % (name or 'default', ' | '.join(line), time.time() - t0), flush=True)
    print('EMU CHECK', 'PASSED' if ok else 'FAILED')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main(big='--big' in sys.argv))
