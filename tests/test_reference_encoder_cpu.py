"""CPU, build container only (needs /root/reference): the product's encoder networks (host code, PyTorch) against the
reference's own classes -- same state-dict keys (checkpoints load unchanged) and same outputs on the same weights
and RNG.  Unavailable third-party packages of the reference (kornia, imageio, soft_renderer, pytorch3d) are stubbed;
they are not used by the compared forward passes."""
import os
import sys
import types

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'model')), reason='reference tree not mounted')


@pytest.fixture(scope='module')
def ref():
    saved_path, saved_cuda = list(sys.path), torch.Tensor.cuda
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, 'third-party'))
    for name in ['soft_renderer', 'pytorch3d', 'pytorch3d.structures', 'pytorch3d.loss', 'pytorch3d.ops',
                 'pytorch3d.ops.knn', 'pytorch3d.structures.pointclouds', 'imageio', 'kornia', 'kornia.geometry', 'trimesh']:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['pytorch3d.ops.knn'].knn_gather = sys.modules['pytorch3d.ops.knn'].knn_points = None
    sys.modules['pytorch3d.structures.pointclouds'].Pointclouds = None
    sys.modules['kornia'].geometry = sys.modules['kornia.geometry']
    sys.modules['kornia.geometry'].quaternion_to_rotation_matrix = lambda q, order=None: torch.eye(3)[None].repeat(q.shape[0], 1, 1)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import torchvision
    real_resnet18 = torchvision.models.resnet18
    torchvision.models.resnet18 = lambda pretrained=False, **k: real_resnet18(weights=None)
    mods = types.SimpleNamespace()
    try:
        from model.module.network import image_encoder, mesh_encoder
        mods.image_encoder, mods.mesh_encoder = image_encoder, mesh_encoder
        try:
            from absl import flags
            from model.module.network import pose_predictor, shape_predictor
            from model.module import mesh as ref_mesh  # noqa: F401  (defines symmetry_idx / init_scale flags)
            mods.pose_predictor, mods.shape_predictor = pose_predictor, shape_predictor
        except Exception as e:      # optional: depends on more third-party imports
            mods.pose_predictor = mods.shape_predictor = None
            mods.err = repr(e)
        yield mods
    finally:
        torchvision.models.resnet18 = real_resnet18
        torch.Tensor.cuda = saved_cuda
        sys.path[:] = saved_path


def test_image_encoder_decoder_match_reference(ref):
    from self_corr_pose_b200.model.module.network.encoder_nets import ResNetEncoder, ResNetDecoder
    torch.manual_seed(0)
    enc, dec = ResNetEncoder().eval(), ResNetDecoder(True, 64, 4).eval()
    r_enc, r_dec = ref.image_encoder.ResNet_Encoder().eval(), ref.image_encoder.ResNet_Decoder(True, 64, 4).eval()
    assert set(enc.state_dict().keys()) == set(r_enc.state_dict().keys())
    assert set(dec.state_dict().keys()) == set(r_dec.state_dict().keys())
    r_enc.load_state_dict(enc.state_dict())
    r_dec.load_state_dict(dec.state_dict())
    x = torch.rand(2, 3, 128, 128)
    with torch.no_grad():
        a, b = dec(*enc(x)), r_dec(*r_enc(x))
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    assert a.shape == (2, 64, 32, 32)


def test_mesh_encoder_matches_reference(ref):
    from self_corr_pose_b200.model.module.network.encoder_nets import MeshEncoder
    torch.manual_seed(1)
    m, r = MeshEncoder(64), ref.mesh_encoder.MeshEncoder(64)
    assert set(m.state_dict().keys()) == set(r.state_dict().keys())
    r.load_state_dict(m.state_dict())
    v = torch.randn(3, 50, 3)
    torch.testing.assert_close(m(v), r(v.clone()), rtol=1e-5, atol=1e-6)


def test_pose_and_shape_heads_match_reference(ref):
    if ref.pose_predictor is None:
        pytest.skip('reference heads not importable here: ' + getattr(ref, 'err', ''))
    from types import SimpleNamespace
    from self_corr_pose_b200.model.module.network.encoder_nets import PosePredictor, ShapePredictor
    opts = SimpleNamespace(depth_offset=5., use_scale=False, symmetry_idx=1, rotation_offset=[0.2, 0, 0, 0, -0.2, 0.2],
                           num_multipose_az=1, num_multipose_el=1, initial_quat_bias_deg=0, baseQuat_elevationBias=0,
                           baseQuat_azimuthBias=0, codedim=64, no_deform=False, deform_ratio=1.)
    torch.manual_seed(2)
    p, r = PosePredictor(opts, 512), ref.pose_predictor.PosePredictor(opts, 512)
    assert set(p.state_dict().keys()) == set(r.state_dict().keys())
    r.load_state_dict(p.state_dict())
    feat = torch.randn(4, 512)
    for a, b in zip(p(feat), r(feat.clone())):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    s, rs = ShapePredictor(opts), ref.shape_predictor.ShapePredictor(opts)
    assert set(s.state_dict().keys()) == set(rs.state_dict().keys())
    rs.load_state_dict(s.state_dict())
    mv, code = torch.randn(2, 40, 3), torch.randn(2, 64)
    torch.testing.assert_close(s(mv, code), rs(mv.clone(), code.clone()), rtol=1e-5, atol=1e-6)


def test_encoder_wrapper_matches_reference(ref):
    """Encoder (model/module/encoder.py:13-52): same submodule / parameter names, same RNG consumption (ColorJitter),
    same six outputs incl. the fp64-promoted principal-point shift of the translation."""
    if ref.pose_predictor is None:
        pytest.skip('reference heads not importable here: ' + getattr(ref, 'err', ''))
    from types import SimpleNamespace
    from model.module.encoder import Encoder as RefEncoder
    from self_corr_pose_b200.model.module.encoder import Encoder
    opts = SimpleNamespace(depth_offset=5., use_scale=False, symmetry_idx=1, rotation_offset=[0.2, 0, 0, 0, -0.2, 0.2],
                           num_multipose_az=1, num_multipose_el=1, initial_quat_bias_deg=0, baseQuat_elevationBias=0,
                           baseQuat_azimuthBias=0, codedim=64, no_deform=False, deform_ratio=1., n_corr_feat=64,
                           img_size=64, corr_h=16, corr_w=16)
    torch.manual_seed(4)
    ours, theirs = Encoder(opts).eval(), RefEncoder(opts).eval()
    assert set(ours.state_dict().keys()) == set(theirs.state_dict().keys())
    theirs.load_state_dict(ours.state_dict())
    g = torch.Generator().manual_seed(5)
    B = 4    # not 3: the reference's Gram-Schmidt calls torch.cross without dim, which picks the batch axis when B == 3
    img = torch.rand(B, 3, 64, 64, generator=g)
    mean_v = torch.randn(B, 30, 3, generator=g) * 0.3
    pp = (0.1 * (torch.rand(B, 2, generator=g) - 0.5)).double()
    foc = (3.7 + 0.3 * torch.rand(B, 2, generator=g)).double()
    outs = []
    for enc in (ours, theirs):
        torch.manual_seed(6)                 # ColorJitter draws its parameters from the global RNG, in eval too
        with torch.no_grad():
            outs.append(enc(img.clone(), mean_v.clone(), pp.clone(), foc.clone()))
    names = ('img_feat', 'mesh_feat', 'pred_v', 'rotation', 'translation', 'scale')
    for n, a, b in zip(names, *outs):
        assert a.shape == b.shape and a.dtype == b.dtype, n
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6, msg=n)
