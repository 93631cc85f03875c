"""GPU: the channels-last glue kernels of the encoder (csrc/scp_nhwc.cu) against the torch operators they replace
(the reference reaches those through torchvision's resnet18 / F.interpolate / F.normalize): values and gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize('shape', [(2, 64, 32, 32), (3, 8, 17, 23), (1, 4, 1, 5), (4, 64, 128, 128)])
def test_maxpool_matches_torch(shape):
    from self_corr_pose_b200.ops import nhwc
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(*shape, device='cuda', generator=g)
    x = torch.relu(x)                      # exact ties at 0, as after the stem's ReLU: the first maximum must win
    x[0, :, 0, 0] = float('nan') if shape[2] > 1 else x[0, :, 0, 0]
    xa, xb = _cl(x).clone().requires_grad_(True), _cl(x).clone().requires_grad_(True)
    ya = nhwc.maxpool3x3s2(xa)
    yb = F.max_pool2d(xb, 3, 2, 1)
    assert ya.shape == yb.shape
    assert torch.equal(torch.nan_to_num(ya, nan=-7.0), torch.nan_to_num(yb, nan=-7.0))
    gy = torch.randn(yb.shape, device='cuda', generator=g)
    ya.backward(_cl(gy)); yb.backward(_cl(gy))
    assert torch.allclose(xa.grad, xb.grad, rtol=1e-6, atol=1e-6)
    assert ya.is_contiguous(memory_format=torch.channels_last) and xa.grad.is_contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize('shape', [(2, 512, 8, 8), (2, 128, 32, 32), (1, 4, 1, 1), (3, 12, 5, 7)])
def test_upsample2x_matches_torch(shape):
    from self_corr_pose_b200.ops import nhwc
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(*shape, device='cuda', generator=g)
    xa, xb = _cl(x).clone().requires_grad_(True), _cl(x).clone().requires_grad_(True)
    ya = nhwc.upsample2x(xa)
    yb = F.interpolate(xb, (2 * shape[2], 2 * shape[3]), mode='bilinear', align_corners=False)
    assert torch.allclose(ya, yb, rtol=0, atol=5e-7), float((ya - yb).abs().max())
    gy = torch.randn(yb.shape, device='cuda', generator=g)
    ya.backward(_cl(gy)); yb.backward(_cl(gy))
    assert torch.allclose(xa.grad, xb.grad, rtol=1e-6, atol=2e-6), float((xa.grad - xb.grad).abs().max())


@pytest.mark.parametrize('shape', [(2, 64, 64, 64), (3, 64, 5, 9), (1, 128, 7, 7), (2, 20, 3, 11)])
def test_l2norm_matches_torch(shape):
    from self_corr_pose_b200.ops import nhwc
    g = torch.Generator(device='cuda').manual_seed(2)
    x = torch.randn(*shape, device='cuda', generator=g) * 3
    x[0, :, 0, 0] = 0                      # zero vector: the eps clamp
    xa, xb = _cl(x).clone().requires_grad_(True), _cl(x).clone().requires_grad_(True)
    ya = nhwc.l2norm_cp(xa)
    yb = F.normalize(xb.flatten(2), p=2, dim=1)
    assert ya.is_contiguous() and ya.shape == yb.shape
    assert torch.allclose(ya, yb, rtol=1e-6, atol=1e-7), float((ya - yb).abs().max())
    gy = torch.randn(yb.shape, device='cuda', generator=g)
    ya.backward(gy); yb.backward(gy)
    ga, gb = xa.grad, xb.grad
    mask = torch.ones_like(ga, dtype=torch.bool)
    mask[0, :, 0, 0] = False               # at the zero vector torch's norm backward is 0 * inf-free but huge (1 / eps): compare apart
    assert torch.allclose(ga[mask], gb[mask], rtol=1e-5, atol=1e-6), float((ga - gb)[mask].abs().max())
    assert torch.isfinite(ga).all()


def test_encoder_uses_the_kernels_and_matches_torch_ops():
    """encode_img with the native glue against the same module evaluated with the torch operators (kernels disabled)."""
    import os
    os.environ.setdefault('SCP_SYNTHETIC_WEIGHTS', '1')
    from self_corr_pose_b200.hotpath import default_opts
    from self_corr_pose_b200.model.module.encoder import Encoder
    from self_corr_pose_b200.ops import nhwc
    torch.manual_seed(0)
    enc = Encoder(default_opts(img_size=128, corr_h=32, corr_w=32)).cuda()
    for m in (enc.backbone, enc.featnet):
        m.to(memory_format=torch.channels_last)
    img = torch.rand(4, 3, 128, 128, device='cuda')
    out = {}
    saved = nhwc.usable
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # fp32 convolutions on both sides: the
    for native in (True, False):                                                      # 4-channel stem picks another cuDNN kernel
        nhwc.usable = saved if native else (lambda t: False)
        try:
            torch.manual_seed(5)
            enc.zero_grad()
            code, feat = enc.encode_img(img)
            (code.square().mean() + (feat * torch.linspace(0, 1, feat.shape[2], device='cuda')).sum() * 1e-3).backward()
            out[native] = (code.detach().clone(), feat.detach().clone(),
                           torch.cat([p.grad.flatten() for p in enc.backbone.parameters() if p.grad is not None]))
        finally:
            nhwc.usable = saved
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    for a, b, tol in zip(out[True], out[False], (1e-5, 1e-5, 1e-4)):
        r = float((a - b).norm() / b.norm().clamp_min(1e-30))
        assert r < tol, r
