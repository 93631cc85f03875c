"""Channels-last glue operators of the convolutional encoder (csrc/scp_nhwc.cu): stem max-pool, the decoder's bilinear 2x
up-sampling and the L2 normalisation of the pixel features, forward + backward, as torch.autograd Functions over NHWC fp32
CUDA tensors.  They replace at::native's NHWC kernels (5-10x below the HBM roofline at the training shape) with the same
arithmetic.  No CPU path: callers keep the torch operators for CPU tensors."""
import torch

from .. import _lib


def _nhwc(t):
    """(B,C,H,W) tensor -> itself if its memory is NHWC-dense, else a channels-last copy."""
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


def usable(t):
    return t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] % 4 == 0


class _MaxPool3x3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _nhwc(x)
        B, C, H, W = x.shape
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty(B, C, OH, OW, device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        idx = torch.empty(B, OH, OW, C, device=x.device, dtype=torch.uint8)
        with torch.cuda.device(x.device):
            rc = _lib.lib().scp_nhwc_maxpool3x3s2_forward(x.data_ptr(), y.data_ptr(), idx.data_ptr(), B, H, W, C,
                                                          _lib.stream_ptr(x.device))
        _lib.check(rc, 'scp_nhwc_maxpool3x3s2_forward')
        ctx.save_for_backward(idx)
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        idx, = ctx.saved_tensors
        B, C, H, W = ctx.shape
        gy = _nhwc(gy)
        gx = torch.empty(B, C, H, W, device=gy.device, dtype=gy.dtype, memory_format=torch.channels_last)
        with torch.cuda.device(gy.device):
            rc = _lib.lib().scp_nhwc_maxpool3x3s2_backward(gy.data_ptr(), idx.data_ptr(), gx.data_ptr(), B, H, W, C,
                                                           _lib.stream_ptr(gy.device))
        _lib.check(rc, 'scp_nhwc_maxpool3x3s2_backward')
        return gx


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _nhwc(x)
        B, C, H, W = x.shape
        y = torch.empty(B, C, 2 * H, 2 * W, device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        with torch.cuda.device(x.device):
            rc = _lib.lib().scp_nhwc_upsample_bilinear_forward(x.data_ptr(), y.data_ptr(), B, H, W, C, 2 * H, 2 * W,
                                                               _lib.stream_ptr(x.device))
        _lib.check(rc, 'scp_nhwc_upsample_bilinear_forward')
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        B, C, H, W = ctx.shape
        gy = _nhwc(gy)
        gx = torch.empty(B, C, H, W, device=gy.device, dtype=gy.dtype, memory_format=torch.channels_last)
        with torch.cuda.device(gy.device):
            rc = _lib.lib().scp_nhwc_upsample2x_bilinear_backward(gy.data_ptr(), gx.data_ptr(), B, H, W, C,
                                                                  _lib.stream_ptr(gy.device))
        _lib.check(rc, 'scp_nhwc_upsample2x_bilinear_backward')
        return gx


class _L2NormCP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        x = _nhwc(x)
        B, C, H, W = x.shape
        P = H * W
        y = torch.empty(B, C, P, device=x.device, dtype=x.dtype)
        inv = torch.empty(B, P, device=x.device, dtype=x.dtype)
        with torch.cuda.device(x.device):
            rc = _lib.lib().scp_nhwc_l2norm_forward(x.data_ptr(), y.data_ptr(), inv.data_ptr(), B, P, C, float(eps),
                                                    _lib.stream_ptr(x.device))
        _lib.check(rc, 'scp_nhwc_l2norm_forward')
        ctx.save_for_backward(y, inv)
        ctx.shape = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        y, inv = ctx.saved_tensors
        B, C, H, W = ctx.shape
        gy = gy.contiguous()
        gx = torch.empty(B, C, H, W, device=gy.device, dtype=gy.dtype, memory_format=torch.channels_last)
        with torch.cuda.device(gy.device):
            rc = _lib.lib().scp_nhwc_l2norm_backward(gy.data_ptr(), y.data_ptr(), inv.data_ptr(), gx.data_ptr(), B, H * W, C,
                                                     _lib.stream_ptr(gy.device))
        _lib.check(rc, 'scp_nhwc_l2norm_backward')
        return gx, None


def maxpool3x3s2(x):
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on an NHWC CUDA tensor."""
    return _MaxPool3x3s2.apply(x)


def upsample2x(x):
    """F.interpolate(x, (2H, 2W), mode='bilinear', align_corners=False) on an NHWC CUDA tensor."""
    return _Upsample2x.apply(x)


def l2norm_cp(x, eps=1e-12):
    """F.normalize(x.flatten(2), p=2, dim=1) for an NHWC CUDA feature map (B,C,H,W) with C <= 128: returns the contiguous
    (B, C, H*W) matrix of unit vectors."""
    return _L2NormCP.apply(x, eps)
