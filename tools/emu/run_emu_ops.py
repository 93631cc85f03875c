"""Functional check WITHOUT a GPU of the fused loss / geometry / cycle / correspondence / encoder-glue / data-path kernels: the
shipped csrc/scp_loss.cu, scp_geom.cu, scp_cycle.cu, scp_corr.cu, scp_nhwc.cu and scp_data.cu compiled for the host (tools/emu/build_emu.py) and called through their C ABI with host
pointers, against the reference's op-by-op statements evaluated in fp64 with torch autograd (the same references the
-m gpu tests use).  Values and gradients.      python tools/emu/run_emu_ops.py"""
import ctypes
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import build_emu  # noqa: E402
from self_corr_pose_b200 import _lib, synthetic  # noqa: E402
from self_corr_pose_b200.model.util import loss_utils as L  # noqa: E402
from self_corr_pose_b200.ops.project_faces import FaceTopology, LOOK_AT_Z  # noqa: E402
from self_corr_pose_b200.soft_renderer import functional as srf  # noqa: E402


def load(source, names):
    lib = ctypes.CDLL(build_emu.build(source=source))
    for n in names:
        fn = getattr(lib, n)
        fn.argtypes, fn.restype = _lib._SIGNATURES[n]
    return lib


def ptr(t):
    return None if t is None else t.data_ptr()


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def views(ts):
    p = (ctypes.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])
    s = (ctypes.c_longlong * len(ts))(*[0 if t is None else t.stride(0) for t in ts])
    return p, s


def check_image_losses(B=2, H=32, hf=8, seed=0):
    lib = load('scp_loss', ('scp_image_losses_workspace_bytes', 'scp_image_losses_forward', 'scp_image_losses_backward'))
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    r_depth = torch.cat([r(B, 2, H, H), 3 * r(B, 1, H, H), (r(B, 1, H, H) - 0.3).clamp(0, 1)], 1)
    r_tex = r(B, 4, H, H)
    r_nocs = torch.cat([r(B, 3, H, H) - 0.5, (r(B, 1, H, H) - 0.4).clamp(0, 1)], 1)
    match_lr = r(B, hf * hf, 3) - 0.5
    img, mask = r(B, 3, H, H), (r(B, H, H) > 0.4).float()
    depth = 3 * r(B, H, H) * (r(B, H, H) > 0.2).float()
    w = r(B, 4) + 0.5
    # reference statements, fp64
    lv = [t.clone().double().requires_grad_(True) for t in (r_depth, r_tex, match_lr)]
    match = F.interpolate(lv[2].reshape(B, hf, hf, 3).permute(0, 3, 1, 2), (H, H), mode='nearest')
    ref = torch.stack([L.compute_mask_loss(img.double(), mask.double(), lv[0][:, 3]),
                       L.compute_texture_loss(img.double(), mask.double(), lv[1][:, :3], lv[1][:, 3]),
                       L.compute_depth_loss(depth.double(), lv[0][:, 2], lv[0][:, 3], mask.double())[0],
                       L.compute_match_loss(match, r_nocs[:, :3].double(), r_nocs[:, 3].double(), mask.double())], 1)
    (ref * w.double()).sum().backward()
    # emulated kernels
    maps = [img, mask, depth, r_depth[:, 3], r_tex[:, :3], r_tex[:, 3], r_depth[:, 2], r_depth[:, 3], r_nocs[:, :3],
            r_nocs[:, 3], None]
    losses = torch.empty(B, 4)
    ws = torch.zeros(lib.scp_image_losses_workspace_bytes(B), dtype=torch.uint8)
    mp, ms = views(maps)
    assert lib.scp_image_losses_forward(mp, ms, ptr(match_lr), B, H, H, hf, hf, 1, ptr(losses), ptr(ws), None) == 0
    g_depth, g_tex, g_match = torch.zeros_like(r_depth), torch.empty_like(r_tex), torch.empty_like(match_lr)
    gp, gs = views([g_depth[:, 3], g_tex[:, :3], g_tex[:, 3], g_depth[:, 2], None])
    assert lib.scp_image_losses_backward(mp, ms, ptr(match_lr), B, H, H, hf, hf, 1, ptr(w), ptr(ws), gp, gs, ptr(g_match),
                                         None) == 0
    return dict(losses=rel(losses, ref), g_r_depth=rel(g_depth, lv[0].grad), g_r_tex=rel(g_tex, lv[1].grad),
                g_match_lr=rel(g_match, lv[2].grad))


def check_geometry(B=3, seed=1):
    lib = load('scp_geom', ('scp_project_faces_forward', 'scp_project_faces_backward', 'scp_spmm3'))
    g = torch.Generator().manual_seed(seed)
    v, f = synthetic.icosphere(1)
    N, nf = v.shape[0], f.shape[0]
    pv = torch.from_numpy(v)[None] + 0.01 * torch.randn(B, N, 3, generator=g)
    rot, trans = (t.contiguous() for t in synthetic.random_poses(B, g))
    foc = (3.7 + 0.3 * torch.rand(B, 2, generator=g)).double()
    pp = (0.1 * (torch.rand(B, 2, generator=g) - 0.5)).double()
    faces = torch.from_numpy(f)
    topo = FaceTopology(faces, N)
    w_sv, w_fv, w_ft = (torch.randn(s, generator=g) for s in ((B, N, 3), (B, nf, 3, 3), (B, nf, 3, 3)))
    lv = [t.clone().double().requires_grad_(True) for t in (pv, rot, trans)]
    sv_r = L.project_to_screen(lv[0], foc, pp, lv[1], lv[2])
    fb = faces[None].repeat(B, 1, 1)
    fv_r = srf.face_vertices(sv_r + torch.tensor([0., 0., LOOK_AT_Z], dtype=torch.float64), fb)
    ft_r = srf.face_vertices(sv_r, fb)
    ((sv_r * w_sv).sum() + (fv_r * w_fv).sum() + (ft_r * w_ft).sum()).backward()
    sv, fv, ft = torch.empty(B, N, 3), torch.empty(B, nf, 3, 3), torch.empty(B, nf, 3, 3)
    t3 = trans.reshape(B, 3).contiguous()
    assert lib.scp_project_faces_forward(ptr(pv), ptr(rot), ptr(t3), ptr(foc), ptr(pp), ptr(topo.faces), B, N, nf,
                                         float(LOOK_AT_Z), ptr(sv), ptr(fv), ptr(ft), None) == 0
    g_v, g_R, g_t = torch.empty(B, N, 3), torch.empty(B, 3, 3), torch.empty(B, 3)
    assert lib.scp_project_faces_backward(ptr(pv), ptr(rot), ptr(t3), ptr(foc), ptr(pp), ptr(topo.csr_off),
                                          ptr(topo.csr_idx), B, N, nf, ptr(w_sv), ptr(w_fv), ptr(w_ft), ptr(g_v), ptr(g_R),
                                          ptr(g_t), None) == 0
    out = dict(screen=rel(sv, sv_r), face_v=rel(fv, fv_r), face_t=rel(ft, ft_r), g_verts=rel(g_v, lv[0].grad),
               g_rot=rel(g_R, lv[1].grad), g_trans=rel(g_t, lv[2].grad.reshape(B, 3)))
    # sparse Laplacian product against the dense buffer of LaplacianLoss
    lap = L.LaplacianLoss(torch.from_numpy(v), faces, average=True)
    x, y = torch.randn(B, N, 3, generator=g), torch.empty(B, N, 3)
    assert lib.scp_spmm3(ptr(lap.csr_off), ptr(lap.csr_col), ptr(lap.csr_val), ptr(x), ptr(y), B, N, None) == 0
    out['laplacian'] = rel(y, torch.matmul(lap.laplacian.double(), x.double()))
    return out


def check_cycle_rows(B=4, P4=64, N=90, k=12, seed=2):
    lib = load('scp_cycle', ('scp_cycle_rows_forward', 'scp_cycle_rows_backward'))
    g = torch.Generator().manual_seed(seed)
    NP, tau = 2 * B, 10.0
    pc = torch.rand(B, P4, N, generator=g) * 2 - 1
    pc[:, ::7] -= 25000.
    A = torch.rand(B, 2, N, generator=g) * 2 - 1
    dw = torch.rand(B, N, generator=g)
    src_idx = torch.arange(NP) % B
    tgt_idx = (torch.arange(NP) + 1 + (torch.arange(NP) // B)) % B
    rows = torch.stack([torch.randperm(P4, generator=g)[:k] for _ in range(NP)])
    pts = torch.rand(NP, 2, k, generator=g) * 2 - 1
    mask_k = (torch.rand(NP, k, generator=g) > 0.3).float()
    w = torch.rand(NP, generator=g) + 0.5
    pc_r, A_r = pc.double().requires_grad_(True), A.double().requires_grad_(True)
    dw_src, dw_tgt = dw.index_select(0, src_idx), dw.index_select(0, tgt_idx)
    A_src = A_r.index_select(0, src_idx) * (dw_src[:, None] >= 0.5)
    s_src = (dw_src >= 0.5).double()
    flat = (tgt_idx[:, None] * P4 + rows).reshape(-1)
    rr = pc_r.reshape(-1, N).index_select(0, flat).reshape(NP, k, N)
    Pi = torch.softmax(tau * rr, dim=2) * (dw_tgt[:, None] >= 0.5)
    match_r = torch.matmul(A_src, Pi.permute(0, 2, 1)) / (torch.matmul(s_src[:, None], Pi.permute(0, 2, 1)) + 1e-5)
    pair_r = ((match_r - pts.double()).norm(2, 1) * mask_k.double()).sum(1)
    (pair_r * w.double()).sum().backward()
    pair, match = torch.empty(NP), torch.empty(NP, 2, k)
    a = (ptr(pc), ptr(A), ptr(dw), ptr(src_idx), ptr(tgt_idx), ptr(rows), ptr(pts), ptr(mask_k), tau, B, P4, N, NP, k)
    assert lib.scp_cycle_rows_forward(*a, ptr(pair), ptr(match), None) == 0
    g_pc, g_A = torch.empty_like(pc), torch.empty_like(A)
    assert lib.scp_cycle_rows_backward(*a, ptr(w), ptr(g_pc), ptr(g_A), None) == 0
    return dict(pair_loss=rel(pair, pair_r), match=rel(match, match_r), g_pointcorr=rel(g_pc, pc_r.grad),
                g_A=rel(g_A, A_r.grad))


def check_correspondence(B=2, hf=16, wf=16, N=70, C=64, seed=0, masks='disc'):
    """scp_corr.cu (mma.sync replaced by its host statement in scp_mma.cuh): forward outputs and both feature gradients,
    fused (row kernel reduces g_mesh_feat too) and split backward, against the reference formulation in fp64."""
    from oracle import corr as ocorr
    lib = load('scp_corr', ('scp_corr_workspace_bytes', 'scp_corr_match_forward', 'scp_corr_match_backward'))
    g = torch.Generator().manual_seed(seed)
    P = hf * wf
    img_feat = F.normalize(torch.randn(B, C, P, generator=g), 2, 1)
    mesh_feat = F.normalize(torch.relu(torch.randn(B, N, C, generator=g)), 2, -1)
    H = 4 * hf
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, H), indexing='ij')
    mask = torch.stack([(((xx - 0.1 * b) ** 2 + yy ** 2) < 0.6 ** 2).float() for b in range(B)])
    if masks == 'mixed':        # image 1 without any foreground (uniform soft-max fallbacks), image 2 without background
        mask[1] = 0.
        mask[2] = 1.
    pred_v = torch.randn(B, N, 3, generator=g)
    w_match, w_imatch = torch.randn(B, P, 3, generator=g), torch.randn(B, 2, N, generator=g)
    w_pool, w_A = torch.randn(B, P // 4, N, generator=g) * 0.01, torch.randn(B, 2, N, generator=g)
    a64 = img_feat.double().requires_grad_(True)
    m64 = mesh_feat.double().requires_grad_(True)
    pc, _, imatch_r, match_r = ocorr.match(a64, m64, mask.double(), pred_v.double(), hf, wf)
    pool_r = F.interpolate(pc.permute(0, 2, 1).reshape(B, N, hf, wf), (hf // 2, wf // 2), mode='bilinear') \
        .reshape(B, N, -1).permute(0, 2, 1)
    grid2 = F.interpolate(ocorr.meshgrid(hf, wf).reshape(1, 2, hf, wf), (hf // 2, wf // 2), mode='bilinear').reshape(2, -1)
    A_r = torch.matmul(grid2.double()[None], torch.softmax(10.0 * pool_r, dim=1))
    ((match_r * w_match).sum() + (imatch_r * w_imatch).sum() + (pool_r * w_pool).sum() + (A_r * w_A).sum()).backward()

    mask_down = F.interpolate(mask[:, None], (hf, wf), mode='nearest').reshape(B, -1).contiguous()
    grid = ocorr.meshgrid(hf, wf).contiguous()
    f32 = lambda *sh: torch.empty(*sh)
    pc_full, pc_pool, match, imatch = f32(B, P, N), f32(B, P // 4, N), f32(B, P, 3), f32(B, 2, N)
    rsum, csum, A_pool, csum_pool = f32(B, P), f32(B, N), f32(B, 2, N), f32(B, N)
    ws_n = lib.scp_corr_workspace_bytes(B, hf, wf, N)
    ws = torch.zeros(max(ws_n, 16), dtype=torch.uint8)
    common = (ptr(img_feat), ptr(mesh_feat), ptr(mask_down), ptr(pred_v), ptr(grid), 10.0, B, hf, wf, N, C)
    assert lib.scp_corr_match_forward(*common, ptr(pc_full), ptr(pc_pool), ptr(match), ptr(imatch), ptr(rsum), ptr(csum),
                                      ptr(A_pool), ptr(csum_pool), ptr(ws), ws_n, None) == 0
    out = dict(pointcorr=rel(pc_full, pc), pool=rel(pc_pool, pool_r), match=rel(match, match_r), imatch=rel(imatch, imatch_r),
               A_pool=rel(A_pool, A_r))
    # training path (no full-resolution output): scp_corr_tc.cu -- operand preparation, epilogue functors and combining
    # kernels as shipped, the tcgen05 GEMM replaced by its host statement; with and without the pooled outputs
    ws.fill_(255)       # scratch the kernels do not write must not be read: NaN patterns instead of zeros
    t_pool, t_match, t_imatch = torch.full((B, P // 4, N), 7.0), f32(B, P, 3), f32(B, 2, N)
    t_rsum, t_csum, t_A, t_cp = f32(B, P), f32(B, N), f32(B, 2, N), f32(B, N)
    assert lib.scp_corr_match_forward(*common, None, ptr(t_pool), ptr(t_match), ptr(t_imatch), ptr(t_rsum), ptr(t_csum),
                                      ptr(t_A), ptr(t_cp), ptr(ws), ws_n, None) == 0
    out.update(tc_pool=rel(t_pool, pool_r), tc_match=rel(t_match, match_r), tc_imatch=rel(t_imatch, imatch_r),
               tc_A_pool=rel(t_A, A_r), tc_rsum=rel(t_rsum, rsum), tc_csum=rel(t_csum, csum), tc_csum_pool=rel(t_cp, csum_pool))
    n_match, n_imatch, n_rsum, n_csum = f32(B, P, 3), f32(B, 2, N), f32(B, P), f32(B, N)
    assert lib.scp_corr_match_forward(*common, None, None, ptr(n_match), ptr(n_imatch), ptr(n_rsum), ptr(n_csum),
                                      None, None, ptr(ws), ws_n, None) == 0
    out.update(tc_nopool_match=rel(n_match, match_r), tc_nopool_imatch=rel(n_imatch, imatch_r))
    os.environ['SCP_CORR_FWD'] = 'legacy'        # the mma.sync kernel on the same call
    l_pool, l_match = f32(B, P // 4, N), f32(B, P, 3)
    assert lib.scp_corr_match_forward(*common, None, ptr(l_pool), ptr(l_match), ptr(imatch), ptr(rsum), ptr(csum),
                                      ptr(A_pool), ptr(csum_pool), ptr(ws), ws_n, None) == 0
    os.environ.pop('SCP_CORR_FWD')
    out.update(tc_vs_legacy_pool=rel(t_pool, l_pool), tc_vs_legacy_match=rel(t_match, l_match))
    for mode in ('fused', 'split'):
        os.environ['SCP_CORR_BWD'] = mode
        g_img, g_mesh = f32(B, C, P), f32(B, N, C)
        assert lib.scp_corr_match_backward(*common, ptr(match), ptr(imatch), ptr(rsum), ptr(csum), ptr(w_match),
                                           ptr(w_imatch), ptr(w_pool), None, ptr(A_pool), ptr(csum_pool), ptr(w_A),
                                           ptr(g_img), ptr(g_mesh), ptr(ws), ws_n, None) == 0
        out['g_img_' + mode], out['g_mesh_' + mode] = rel(g_img, a64.grad), rel(g_mesh, m64.grad)
    os.environ.pop('SCP_CORR_BWD', None)
    return out


def check_nhwc(seed=3):
    """scp_nhwc.cu: max-pool, bilinear 2x, channel L2 norm (forward + backward) against the torch operators on CPU."""
    names = ('scp_nhwc_maxpool3x3s2_forward', 'scp_nhwc_maxpool3x3s2_backward', 'scp_nhwc_upsample_bilinear_forward',
             'scp_nhwc_upsample2x_bilinear_backward', 'scp_nhwc_l2norm_forward', 'scp_nhwc_l2norm_backward')
    lib = load('scp_nhwc', names)
    g = torch.Generator().manual_seed(seed)
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    out = {}
    # max-pool (exact ties at 0 as after a ReLU; odd sizes)
    B, C, H, W = 2, 8, 9, 7
    x = cl(torch.relu(torch.randn(B, C, H, W, generator=g)))
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y, idx = cl(torch.empty(B, C, OH, OW)), torch.empty(B, OH, OW, C, dtype=torch.uint8)
    assert lib.scp_nhwc_maxpool3x3s2_forward(ptr(x), ptr(y), ptr(idx), B, H, W, C, None) == 0
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    gy = cl(torch.randn(B, C, OH, OW, generator=g))
    yr.backward(gy)
    gx = cl(torch.empty(B, C, H, W))
    assert lib.scp_nhwc_maxpool3x3s2_backward(ptr(gy), ptr(idx), ptr(gx), B, H, W, C, None) == 0
    out['pool'], out['pool_g'] = float((y - yr).abs().max()), rel(gx, xr.grad)
    # bilinear 2x
    B, C, H, W = 2, 12, 5, 6
    x = cl(torch.randn(B, C, H, W, generator=g))
    y = cl(torch.empty(B, C, 2 * H, 2 * W))
    assert lib.scp_nhwc_upsample_bilinear_forward(ptr(x), ptr(y), B, H, W, C, 2 * H, 2 * W, None) == 0
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, (2 * H, 2 * W), mode='bilinear', align_corners=False)
    gy = cl(torch.randn(B, C, 2 * H, 2 * W, generator=g))
    yr.backward(gy)
    gx = cl(torch.empty(B, C, H, W))
    assert lib.scp_nhwc_upsample2x_bilinear_backward(ptr(gy), ptr(gx), B, H, W, C, None) == 0
    out['up'], out['up_g'] = rel(y, yr), rel(gx, xr.grad)
    # L2 norm over channels: NHWC in -> (B, C, P) out
    B, C, H, W = 2, 64, 5, 9
    x = cl(torch.randn(B, C, H, W, generator=g) * 3)
    P = H * W
    y, inv = torch.empty(B, C, P), torch.empty(B, P)
    assert lib.scp_nhwc_l2norm_forward(ptr(x), ptr(y), ptr(inv), B, P, C, 1e-12, None) == 0
    xr = x.clone().requires_grad_(True)
    yr = F.normalize(xr.flatten(2), p=2, dim=1)
    gy = torch.randn(B, C, P, generator=g)
    yr.backward(gy)
    gx = cl(torch.empty(B, C, H, W))
    assert lib.scp_nhwc_l2norm_backward(ptr(gy), ptr(y), ptr(inv), ptr(gx), B, P, C, None) == 0
    out['l2'], out['l2_g'] = rel(y, yr), rel(gx, xr.grad)
    return out


def check_data(seed=4):
    """scp_data.cu: crop box + resized crops against the reference's dataset class (golden) -- mismatching elements."""
    import numpy as np
    lib = load('scp_data', ('scp_data_bbox_crop', 'scp_data_resized_crop'))
    G = np.load(os.path.join(ROOT, 'tests', 'golden', 'data_golden.npz'))
    n, S = G['raw_img'].shape[0], 64
    img, mask = torch.from_numpy(G['raw_img']).contiguous(), torch.from_numpy(G['raw_mask']).contiguous()
    depth = torch.from_numpy(G['raw_depth'].astype(np.uint16).view(np.int16)).contiguous()
    H, W = mask.shape[1:]
    K = G['K']
    intr = torch.from_numpy(np.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]], 1)).contiguous()
    out = {}
    for tag, no_stretch, aa in (('stretch_noaa', 0, 0), ('nostretch_aa', 1, 1)):
        rs = torch.from_numpy(G[tag + '_rand_scale']).contiguous()
        crop, status = torch.empty(n, 4, dtype=torch.int32), torch.empty(n, dtype=torch.int32)
        center, length = torch.empty(n, 2, dtype=torch.int64), torch.empty(n, 2, dtype=torch.int64)
        fc, pc = torch.empty(n, 2, dtype=torch.float64), torch.empty(n, 2, dtype=torch.float64)
        assert lib.scp_data_bbox_crop(ptr(mask), ptr(rs), ptr(intr), n, H, W, S, no_stretch, ptr(crop), ptr(center), ptr(length),
                                      ptr(fc), ptr(pc), ptr(status), None) == 0
        o_img, o_mask, o_depth = torch.empty(n, 3, S, S), torch.empty(n, 1, S, S), torch.empty(n, 1, S, S)
        assert lib.scp_data_resized_crop(ptr(img), ptr(mask), ptr(depth), ptr(crop), n, H, W, S, 1, aa, ptr(o_img), ptr(o_mask),
                                         ptr(o_depth), None) == 0
        bad = int((center.numpy() != G[tag + '_center']).sum() + (length.numpy() != G[tag + '_length']).sum() +
                  (o_mask.numpy() != G[tag + '_mask']).sum() + (o_depth.numpy() != G[tag + '_depth']).sum() +
                  (o_img.numpy() != G[tag + '_img']).sum() + int(status.sum()))
        out[tag + '_mismatches'] = float(bad)
        out[tag + '_intr'] = max(float(np.abs(fc.numpy() - G[tag + '_foc_crop']).max()), float(np.abs(pc.numpy() - G[tag + '_pp_crop']).max()))
    return out


def check_posefit(L=3, n=1300, H=100, seed=5):
    """scp_posefit.cu: residual table and inlier moments against the host formulation (model/util/umeyama.py) on CPU."""
    from self_corr_pose_b200.model.util import umeyama as U
    lib = load('scp_posefit', ('scp_posefit_chunks', 'scp_posefit_residual_table', 'scp_posefit_inlier_moments'))
    g = torch.Generator().manual_seed(seed)
    src = torch.rand(L, n, 3, generator=g) - 0.5
    R0 = torch.linalg.qr(torch.randn(L, 3, 3, generator=g))[0]
    tgt = 300 * src @ R0 + torch.tensor([0., 0., 900.]) + 3 * torch.randn(L, n, 3, generator=g)
    tgt[:, ::7] += 80 * torch.randn(L, (n + 6) // 7, 3, generator=g)          # outliers
    counts = torch.tensor([n, n - 401, 37][:L], dtype=torch.int32)
    valid = torch.arange(n)[None] < counts[:, None]
    idx = torch.stack([torch.randint(0, int(c), (H, 5), generator=g) for c in counts])
    rows = torch.arange(L)[:, None, None]
    hs, hR, ht, _ = U._closed_form(src[rows, idx], tgt[rows, idx], strict=False)
    pr = U._point_residuals(hs, hR, ht, src, tgt)
    want = torch.linalg.norm(torch.where(valid[:, None], pr, 0 * pr), dim=-1)
    A = (hs[..., None, None] * hR).contiguous()
    nchunk = lib.scp_posefit_chunks(n)
    partial = torch.full((L, nchunk, H), float('nan'))
    ht = ht.contiguous()
    assert lib.scp_posefit_residual_table(ptr(src), ptr(tgt), ptr(counts), ptr(A), ptr(ht), L, n, H, ptr(partial), None) == 0
    out = dict(table=rel(partial.sum(1).sqrt(), want))
    best = want.argmin(-1)
    pick = lambda x: x[torch.arange(L), best]
    pass_t, _ = U._thresholds(src, tgt, valid)
    found = torch.tensor([1, 0, 1][:L], dtype=torch.uint8)                      # image 1: rejected -> all real points
    prb = U._point_residuals(pick(hs)[:, None], pick(hR)[:, None], pick(ht)[:, None], src, tgt)[:, 0]
    inl = (prb < pass_t[:, None]) & valid
    okk = found.bool() & (inl.sum(-1).float() / counts.float() >= 0.1)
    safe = torch.where(okk[:, None], inl, valid)
    fs, fR, ft, _ = U._closed_form(src.double(), tgt.double(), safe, strict=False)
    mom = torch.full((L, 18), float('nan'))
    bA, bt, pt = pick(A).contiguous(), pick(ht).contiguous(), pass_t.contiguous()      # keep the buffers alive over the call
    assert lib.scp_posefit_inlier_moments(ptr(src), ptr(tgt), ptr(counts), ptr(bA), ptr(bt), ptr(pt), ptr(found), L, n, ptr(mom),
                                          None) == 0
    gs, gR, gt, _ = U._from_moments(mom[:, 2:5], mom[:, 5:8], mom[:, 8:17].reshape(L, 3, 3) / mom[:, 0, None, None],
                                    mom[:, 17] / (mom[:, 0] - 1), strict=False)
    out.update(n_used=float((mom[:, 0] - safe.sum(-1)).abs().max()), n_inl=float((mom[:, 1] - inl.sum(-1)).abs().max()),
               scale=rel(gs, fs), rotation=rel(gR, fR), translation=rel(gt, ft))
    return out


def main():
    ok = True
    for name, fn, tol in (('image losses', check_image_losses, 1e-5), ('geometry', check_geometry, 1e-5),
                          ('cycle rows', check_cycle_rows, 1e-4), ('correspondence', check_correspondence, 1e-3),
                          ('corr. masks', lambda: check_correspondence(B=3, hf=16, wf=32, N=200, seed=2, masks='mixed'), 1e-3),
                          ('nhwc glue', check_nhwc, 2e-6), ('data path', check_data, 1e-12),
                          ('pose fit', check_posefit, 2e-5)):
        res = fn()
        ok &= all(v <= tol for v in res.values())
        print('%-13s %s' % (name, '  '.join('%s %.1e' % kv for kv in res.items())), flush=True)
    print('EMU OPS CHECK', 'PASSED' if ok else 'FAILED')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
