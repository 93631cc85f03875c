"""CPU: host-side logic -- C-ABI exports, loss/camera helpers against the reference statements, data-parallel
gradient reduction over gloo (world size 2), batch sharding."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from self_corr_pose_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, 'include', 'scp_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    names = sorted(set(re.findall(r'\b(scp_[a-z0-9_]+)\s*\(', header)))
    assert len(names) >= 10, names
    for n in names:
        assert hasattr(lib, n), 'libscp_b200.so does not export %s' % n
    lib.scp_abi_version.restype = ctypes.c_int
    assert lib.scp_abi_version() == 7
    # size queries are pure host code
    lib.scp_softras_workspace_bytes.restype = ctypes.c_size_t
    assert lib.scp_softras_workspace_bytes(2, 100) >= 2 * 100 * (16 + 192)
    lib.scp_vit_workspace_bytes.restype = ctypes.c_size_t
    assert lib.scp_vit_workspace_bytes(1, 256, 256, 0) > 1025 * 384 * 4
    assert lib.scp_vit_workspace_bytes(1, 250, 256, 0) == 0
    assert lib.scp_vit_workspace_bytes(1, 256, 256, 1) > lib.scp_vit_workspace_bytes(1, 256, 256, 0)   # split pairs


def test_product_fails_loudly_without_cuda():
    from self_corr_pose_b200.soft_renderer import functional as srf
    from self_corr_pose_b200.ops.corr_match import corr_match
    from self_corr_pose_b200.model.module.network.dino import DINO
    with pytest.raises(TypeError):
        srf.soft_rasterize(torch.zeros(1, 1, 3, 3), torch.zeros(1, 1, 1, 3))
    with pytest.raises(TypeError):
        corr_match(torch.zeros(1, 64, 128), torch.zeros(1, 8, 64), torch.zeros(1, 128), torch.zeros(1, 8, 3),
                   torch.zeros(2, 128), 10., 8, 16)
    if not torch.cuda.is_available():
        with pytest.raises(TypeError):
            DINO()(torch.zeros(1, 3, 64, 64))


def test_loss_helpers_match_reference_statements():
    """mask pyramid pools along rows only; divide_by_* pairing; pinhole camera in fp64 intrinsics."""
    from self_corr_pose_b200.model.util import loss_utils as L
    from oracle import corr as ocorr
    g = torch.Generator().manual_seed(0)
    x = torch.rand(8, 3, 5, generator=g)
    for name in ('frame', 'instance', 'both'):
        a = getattr(L, 'divide_by_' + name)(x, 2, 4)
        b = ocorr.DIVIDE[name](x, 2, 4)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    m, mp_ = torch.rand(2, 32, 32, generator=g), torch.rand(2, 32, 32, generator=g)
    got = L.compute_mask_loss(None, m, mp_)
    want = 0
    for i in range(5):   # explicit 1-D pooling along the last axis
        k = 2 ** i
        d = (mp_.reshape(2, 32, 32 // k, k).mean(-1) - m.reshape(2, 32, 32 // k, k).mean(-1)).pow(2)
        want = want + d.repeat_interleave(k, dim=2)
    torch.testing.assert_close(got, 0.2 * want.mean((1, 2)), rtol=1e-5, atol=1e-7)
    v = torch.rand(2, 7, 3, generator=g) + torch.tensor([0., 0., 2.])
    foc = torch.tensor([[3.7, 3.6], [3.5, 3.4]], dtype=torch.float64)
    pp = torch.tensor([[0.01, -0.02], [0.0, 0.03]], dtype=torch.float64)
    out = L.pinhole_cam(v.clone(), pp, foc)
    ref = v.clone().double()
    ref[:, :, 0] = pp[:, 0:1] + v[:, :, 0].double() * foc[:, 0:1] / v[:, :, 2].double()
    ref[:, :, 1] = pp[:, 1:2] + v[:, :, 1].double() * foc[:, 1:2] / v[:, :, 2].double()
    torch.testing.assert_close(out, ref.float(), rtol=0, atol=0)


def test_weights_schedule():
    from self_corr_pose_b200.hotpath import default_opts
    from self_corr_pose_b200.model.module.weights import Weights
    w = Weights(default_opts())
    w.schedule(0)
    assert w.triangle_wt == pytest.approx(0.002) and w.match_wt == pytest.approx(0.002)   # match starts at decay*wt
    w.schedule(20000)
    assert w.triangle_wt == pytest.approx(0.0002) and w.match_wt == pytest.approx(0.02)


def test_shard_batch():
    from self_corr_pose_b200.dist import shard_batch
    assert list(shard_batch(256, 4, 3, 8)) == list(range(96, 128))
    with pytest.raises(ValueError):
        shard_batch(30, 4, 0, 8)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from self_corr_pose_b200.dist import FlatGradReducer
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    torch.manual_seed(0)
    w = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2))]
    cl = torch.nn.Parameter(torch.randn(4, 3, 2, 2).to(memory_format=torch.channels_last))   # a channels-last conv weight
    x = torch.full((5, 3), float(rank + 1))
    loss = (w[0] * x).sum() + (w[1] ** 2).sum() * (rank + 1)      # w[2] gets no gradient on any rank
    (loss + (cl * (rank + 1)).sum()).backward()
    red = FlatGradReducer(w + [cl])
    red.reduce()
    assert cl.grad.stride() == cl.stride() and torch.allclose(cl.grad, torch.full_like(cl, 1.5))
    first = [None if p.grad is None else p.grad.clone() for p in w]
    # second step: gradients accumulate in place into the views of the flat buffer
    assert red.adopted and all(p.grad.data_ptr() >= red.flat.data_ptr() for p in w[:2])
    red.zero_()
    assert float(w[0].grad.abs().sum()) == 0
    loss = (w[0] * x).sum() * 3 + (w[1] ** 2).sum() * (rank + 1)
    loss.backward()
    red.reduce()
    out[rank] = first + [p.grad.clone() for p in w[:2]]
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    w0, w1 = torch.randn(5, 3), torch.randn(7)
    exp0 = torch.full((5, 3), 1.5)            # mean of rank-wise gradients 1 and 2
    exp1 = 2 * w1 * 1.5
    for r in range(world):
        torch.testing.assert_close(out[r][0], exp0)
        torch.testing.assert_close(out[r][1], exp1)
        assert out[r][2] is None                  # never reached by the graph: stays None on every rank (as with 1 GPU)
        torch.testing.assert_close(out[r][3], 3 * exp0)
        torch.testing.assert_close(out[r][4], exp1)


def test_face_topology_csr():
    """vertex -> face-corner adjacency of ops/project_faces.py: every corner listed exactly once under its vertex,
    corners of a vertex in ascending order (the fixed summation order of the geometry backward)."""
    from self_corr_pose_b200 import synthetic
    from self_corr_pose_b200.ops.project_faces import FaceTopology
    v, f = synthetic.icosphere(2)
    topo = FaceTopology(torch.from_numpy(f), v.shape[0])
    off, idx, faces = topo.csr_off.long(), topo.csr_idx.long(), topo.faces.long()
    assert off[0] == 0 and off[-1] == faces.numel() and sorted(idx.tolist()) == list(range(faces.numel()))
    flat = faces.reshape(-1)
    for n in range(v.shape[0]):
        seg = idx[off[n]:off[n + 1]]
        assert (flat[seg] == n).all() and (seg[1:] > seg[:-1]).all()


def test_laplacian_csr_equals_dense_buffer():
    from self_corr_pose_b200 import synthetic
    from self_corr_pose_b200.model.util.loss_utils import LaplacianLoss
    v, f = synthetic.icosphere(1)
    lap = LaplacianLoss(torch.from_numpy(v), torch.from_numpy(f), average=True)
    for pre, dense in (('csr', lap.laplacian), ('csr_t', lap.laplacian.t())):
        off, col, val = (getattr(lap, pre + s) for s in ('_off', '_col', '_val'))
        rebuilt = torch.zeros_like(dense)
        for n in range(dense.shape[0]):
            rebuilt[n, col[off[n]:off[n + 1]].long()] = val[off[n]:off[n + 1]]
        assert torch.equal(rebuilt, dense.contiguous())
    assert 'csr_off' not in lap.state_dict()          # derived buffers stay out of checkpoints


def test_argmax_decode():
    from self_corr_pose_b200.model.module.pretrained_corr import decode_argmax

    def pack(x, col):   # what the arg-max epilogue writes (csrc/scp_vit.cu: EpiArgmax)
        u = int(np.float32(x).view(np.uint32))
        key = (~u & 0xffffffff) if u & 0x80000000 else (u | 0x80000000)
        v = (key << 32) | (0xffffffff - col)
        return v - (1 << 64) if v >= (1 << 63) else v      # stored in an int64 tensor

    cands = [(-3.5, 7), (2.25, 900), (2.25, 12), (-0.0, 3), (1e-30, 5)]
    best = max(cands, key=lambda c: (np.float32(c[0]), -c[1]))
    packed = [pack(*c) for c in cands]
    as_u64 = lambda v: v + (1 << 64) if v < 0 else v
    assert max(packed, key=as_u64) == pack(*best)          # larger similarity wins, lower column wins ties
    t = torch.tensor([pack(2.25, 12), 0, pack(-7.0, 1023)], dtype=torch.int64)
    assert decode_argmax(t).tolist() == [12, 0, 1023]


def test_new_ops_fail_loudly_on_cpu_tensors():
    from self_corr_pose_b200.ops.image_losses import image_losses
    from self_corr_pose_b200.ops.project_faces import project_faces
    from self_corr_pose_b200.ops.cycle_rows import cycle_rows
    z = torch.zeros
    with pytest.raises(TypeError):
        image_losses(z(1, 4, 16, 16), z(1, 4, 16, 16), z(1, 16, 3), z(1, 3, 16, 16), z(1, 16, 16), z(1, 16, 16),
                     z(1, 4, 16, 16), 4, 4)
    with pytest.raises(TypeError):
        project_faces(z(1, 4, 3), z(1, 3, 3), z(1, 1, 3), z(1, 2).double(), z(1, 2).double())
    with pytest.raises(TypeError):
        cycle_rows(z(1, 4, 8), z(1, 2, 8), z(1, 8), z(2).long(), z(2).long(), z(2, 3).long(), z(2, 2, 3), z(2, 3), 10.)


def test_buffer_size_validation():
    """The C-ABI takes bare pointers: the wrappers refuse buffers whose element count differs from what the kernel
    indexes, and a face list that points outside the vertex array."""
    from self_corr_pose_b200 import _lib
    from self_corr_pose_b200.ops.project_faces import FaceTopology
    _lib.expect_numel('op', a=(torch.zeros(2, 3), 6), b=(None, 5), c=(torch.zeros(4, 1, 2), 8))
    with pytest.raises(ValueError, match='`a` has 6 elements'):
        _lib.expect_numel('op', a=(torch.zeros(2, 3), 8))
    FaceTopology(torch.tensor([[0, 1, 2], [2, 1, 3]]), 4)
    with pytest.raises(ValueError):
        FaceTopology(torch.tensor([[0, 1, 4]]), 4)
    with pytest.raises(ValueError):
        FaceTopology(torch.tensor([[0, -1, 2]]), 4)


def test_op_shape_checks_accept_the_step_shapes_and_reject_short_buffers():
    """check_shapes of every fused op on the shapes the training step passes (hotpath.py / model.py call sites)."""
    from self_corr_pose_b200.ops import corr_match, cycle_rows, image_losses, project_faces
    z = torch.zeros
    B, H, hf, N, C, nf = 4, 64, 16, 42, 64, 80
    P = hf * hf
    # correspondence: training-step call and the rotation-cycle call (target pixels in the role of vertices)
    assert corr_match.check_shapes(z(B, C, P), z(B, N, C), z(B, P), z(B, N, 3), z(2, P), hf, hf) == (B, C, P, N)
    assert corr_match.check_shapes(z(B, C, P // 4), z(B, P // 4, C), z(B, P // 4), z(B, P // 4, 3), z(2, P // 4),
                                   hf // 2, hf // 2) == (B, C, P // 4, P // 4)
    with pytest.raises(ValueError):
        corr_match.check_shapes(z(B, C, P), z(B, N, C), z(B, P), z(B, N, 3), z(2, P), hf, hf // 2)
    with pytest.raises(ValueError):
        corr_match.check_shapes(z(B, C, P), z(B, N, C), z(B, P // 4), z(B, N, 3), z(2, P), hf, hf)
    # image losses: raw renders + (B,3,H,W) image, (B,H,W) mask / depth, low-resolution match
    args = lambda **kw: dict(dict(r_depth=z(B, 4, H, H), r_tex=z(B, 4, H, H), match_lr=z(B, P, 3), img=z(B, 3, H, H),
                                  mask=z(B, H, H), depth=z(B, H, H), r_nocs=z(B, 4, H, H), hf=hf, wf=hf), **kw)
    assert image_losses.check_shapes(**args()) == (B, H, H)
    assert image_losses.check_shapes(**args(depth=None)) == (B, H, H)
    for bad in (dict(mask=z(B, H, H // 2)), dict(match_lr=z(B, P // 4, 3)), dict(r_tex=z(B, 3, H, H)),
                dict(img=z(B, 1, H, H)), dict(depth=z(B - 1, H, H))):
        with pytest.raises(ValueError):
            image_losses.check_shapes(**args(**bad))
    # geometry
    topo = project_faces.FaceTopology(torch.randint(0, N, (nf, 3)), N)
    good = (z(B, N, 3), z(B, 3, 3), z(B, 1, 3), z(B, 2).double(), z(B, 2).double())
    assert project_faces.check_shapes(*good, topo) == (B, N) and project_faces.check_shapes(*good, None) == (B, N)
    with pytest.raises(ValueError):
        project_faces.check_shapes(z(B, N + 1, 3), *good[1:], topo)
    with pytest.raises(ValueError):
        project_faces.check_shapes(good[0], z(B, 3), *good[2:], None)
    # cycle rows: 2B pairs of k gathered rows
    NP, k = 2 * B, 7
    good = (z(B, P // 4, N), z(B, 2, N), z(B, N), z(NP).long(), z(NP).long(), z(NP, k).long(), z(NP, 2, k), z(NP, k))
    assert cycle_rows.check_shapes(*good) == (B, P // 4, N, NP, k)
    with pytest.raises(ValueError):
        cycle_rows.check_shapes(*good[:6], z(NP, k), z(NP, k))


def test_every_entry_point_rejects_null_arguments_with_a_message():
    """Error behaviour of the C ABI: bad arguments return non-zero BEFORE any CUDA call and leave a text for
    scp_last_error() -- checked for every compute entry point with all-null / all-zero arguments (no GPU involved)."""
    from self_corr_pose_b200 import _lib
    L = _lib.lib()
    checked = 0
    for name, (argtypes, restype) in _lib._SIGNATURES.items():
        if restype is not ctypes.c_int or not argtypes:
            continue
        args = [0 if t in (ctypes.c_int, ctypes.c_size_t) else 0.0 if t is ctypes.c_float else None for t in argtypes]
        assert getattr(L, name)(*args) != 0, name
        msg = L.scp_last_error().decode()
        assert name.replace('scp_', '').split('_')[0] in msg or 'gemm' in msg, (name, msg)
        with pytest.raises(_lib.ScpNativeError, match='failed'):
            _lib.check(-1, name)
        checked += 1
    assert checked >= 17


def test_variant_library_is_opt_in_only():
    """Kernel-experiment variants (build.py --variant, SCP_LIB_VARIANT) never shadow the product library: without the
    variable the product path is loaded, with it a missing variant fails loudly instead of falling back."""
    import subprocess
    import sys
    from self_corr_pose_b200 import _lib, build
    assert os.path.basename(_lib.LIB_PATH) == 'libscp_b200.so' and build.lib_path() == _lib.LIB_PATH
    assert os.path.basename(build.lib_path('x')) == 'libscp_b200.x.so'
    code = ('from self_corr_pose_b200 import _lib\n'
            'try:\n    _lib.lib()\nexcept _lib.ScpNativeError as e:\n    print("LOUD", e)\n')
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=ROOT,
                       env=dict(os.environ, SCP_LIB_VARIANT='does_not_exist'))
    assert 'LOUD' in r.stdout and 'libscp_b200.does_not_exist.so' in r.stdout, r.stdout + r.stderr


def test_unimplemented_reference_options_fail_loudly():
    """Options of the reference's MeshNet that this package does not implement raise instead of being ignored."""
    from self_corr_pose_b200.hotpath import default_opts
    from self_corr_pose_b200.model.model import MeshNet
    for flag in ('flatten_loss', 'depth_loss_chamfer', 'use_occ'):
        opts = default_opts()
        setattr(opts, flag, True)
        with pytest.raises(NotImplementedError, match=flag):
            MeshNet(opts)


# ---- round 2: host pieces of the graphed training step ------------------------------------------------------------------
def test_rotation_slot_equals_torchvision_rotate():
    """RotationSlot (affine matrix through a static tensor, sampling grid from cached constants) reproduces
    torchvision.transforms.functional.rotate bit for bit and draws the angle like correspondence.py:82."""
    import torchvision
    from torchvision.transforms import InterpolationMode
    from self_corr_pose_b200.model.module.correspondence import RotationSlot
    slot = RotationSlot('cpu')
    for shape in ((2, 3, 32, 32), (2, 1, 48, 32), (1, 2, 16, 16)):
        img = torch.rand(*shape)
        torch.manual_seed(3)
        slot.refresh()
        slot.upload()
        after = torch.rand(1)
        torch.manual_seed(3)
        angle = torch.empty(1).uniform_(0., 360.).item()
        assert torch.equal(after, torch.rand(1)) and angle == slot.angle
        for mode, im in (('bilinear', InterpolationMode.BILINEAR), ('nearest', InterpolationMode.NEAREST)):
            want = torchvision.transforms.functional.rotate(img, angle, interpolation=im)
            assert torch.equal(slot.rotate(img, mode), want)


def test_weights_device_buffer_follows_the_schedule():
    from self_corr_pose_b200.hotpath import default_opts
    from self_corr_pose_b200.model.module.weights import Weights
    a, b = Weights(default_opts()), Weights(default_opts())
    b.enable_device_buffer('cpu')
    for it in (0, 1, 777, 19999, 30000):
        a.schedule(it)
        b.schedule(it)
        for name in ('mask_wt', 'tex_wt', 'depth_wt', 'triangle_wt', 'symmetry_wt', 'cycle_loss_wt', 'cycle_loss_pt_wt',
                     'match_wt', 'imatch_wt', 'pullfar_wt', 'deform_wt'):
            assert float(getattr(b, name)) == pytest.approx(float(getattr(a, name)), rel=1e-7), (it, name)
        assert getattr(b, 'match_wt').dim() == 0 and getattr(b, 'match_wt').data_ptr() != 0     # a view of the static buffer


def test_jitter_slot_packs_what_get_params_draws():
    import struct
    from torchvision import transforms
    from self_corr_pose_b200.ops.color_jitter import JitterSlot, _pack
    jit = transforms.ColorJitter(0.2, 0.2, 0.2, 0.05)
    slot = JitterSlot('cpu', (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    torch.manual_seed(9)
    slot.refresh(jit)
    after = torch.rand(1)
    torch.manual_seed(9)
    params = transforms.ColorJitter.get_params(jit.brightness, jit.contrast, jit.saturation, jit.hue)
    assert torch.equal(after, torch.rand(1))                       # same consumption of the global CPU generator
    order, ratios, hue = _pack(params, None, None)
    got = struct.unpack(JitterSlot.FMT, bytes(slot.host.tolist()))
    assert list(got[:4]) == order
    assert got[4:7] == pytest.approx(ratios[0::2]) and got[7:10] == pytest.approx(ratios[1::2]) and got[10] == pytest.approx(hue)
    assert got[11:14] == pytest.approx((0.485, 0.456, 0.406)) and got[14:17] == pytest.approx((0.229, 0.224, 0.225))


def test_surface_sampling_is_area_weighted_and_capturable_ops_only():
    """Inverse-CDF face draws (no torch.multinomial): indices in range, frequencies follow the face areas, barycentric
    weights are a partition of unity."""
    from self_corr_pose_b200 import synthetic
    from self_corr_pose_b200.model.module.mesh import sample_faces_and_weights, points_from_samples
    v, f = synthetic.icosphere(1)
    verts = torch.from_numpy(v)[None].clone()
    verts[0, :, 0] *= 3.0                                            # unequal face areas
    faces = torch.from_numpy(f)[None]
    torch.manual_seed(0)
    idx, w = sample_faces_and_weights(verts, faces, 200000)
    assert idx.min() >= 0 and idx.max() < faces.shape[1] and torch.allclose(w.sum(-1), torch.ones(1, 200000), atol=1e-6)
    assert (w >= 0).all()
    tri = verts[0][faces[0]]
    area = 0.5 * torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=-1).norm(dim=-1)
    freq = torch.bincount(idx[0], minlength=faces.shape[1]).float() / idx.shape[1]
    assert torch.allclose(freq, area / area.sum(), atol=3e-3)
    pts = points_from_samples(verts, faces, idx[:, :1000], w[:, :1000])
    assert pts.shape == (1, 1000, 3) and torch.isfinite(pts).all()


def test_split_bf16_pairs_carry_sixteen_mantissa_bits():
    from self_corr_pose_b200.model.module.network.dino import split_bf16_i32, merge_bf16_i32
    x = torch.randn(5, 96, generator=torch.Generator().manual_seed(1)) * torch.logspace(-3, 3, 96)
    s = split_bf16_i32(x)
    assert s.shape == (5, 192) and s.dtype == torch.bfloat16
    # i32 layout: groups of 32 columns stored [32 hi | 32 lo]
    assert torch.equal(s[:, :32].float(), x[:, :32].to(torch.bfloat16).float())
    assert torch.equal(s[:, 64:96].float(), x[:, 32:64].to(torch.bfloat16).float())
    rel = ((merge_bf16_i32(s) - x).abs() / x.abs()).max()
    assert float(rel) < 2.0 ** -16
