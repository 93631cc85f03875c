"""ctypes binding of the C-ABI library (include/scp_b200.h -> libscp_b200.so).

This is the only place the product touches native code.  There is NO fallback: if the library
is missing or a call fails, an exception is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libscp_b200.so')
# kernel experiments only (build.py --variant): SCP_LIB_VARIANT=NAME loads libscp_b200.NAME.so instead
_VARIANT = os.environ.get('SCP_LIB_VARIANT', '')
if _VARIANT:
    LIB_PATH = os.path.join(_HERE, 'libscp_b200.%s.so' % _VARIANT)
ABI_VERSION = 7
VIT_BF16, VIT_X3 = 0, 1      # `precision` of the ViT entry points (include/scp_b200.h)

_f = ctypes.c_void_p   # device pointers travel as integers
_i = ctypes.c_int
_fl = ctypes.c_float
_sz = ctypes.c_size_t

_SOFTRAS_SCALARS = [_i, _i, _i, _i, _fl, _fl, _fl, _fl, _i, _fl, _fl, _i, _i, _i, _i]

_SIGNATURES = {
    'scp_abi_version': ([], _i),
    'scp_last_error': ([], ctypes.c_char_p),
    'scp_softras_workspace_bytes': ([_i, _i], _sz),
    'scp_softras_forward': ([_f] * 5 + _SOFTRAS_SCALARS + [_f, _sz, _f], _i),
    'scp_softras_backward': ([_f] * 8 + _SOFTRAS_SCALARS + [_f, _sz, _f], _i),
    'scp_softras_forward_dual': ([_f] * 8 + [_i, _i, _i, _fl, _fl, _fl, _fl, _fl, _fl, _i, _f, _sz, _f], _i),
    'scp_corr_workspace_bytes': ([_i, _i, _i, _i], _sz),
    'scp_corr_backward_workspace_bytes': ([_i, _i, _i, _i], _sz),
    'scp_corr_match_forward': ([_f] * 5 + [_fl, _i, _i, _i, _i, _i] + [_f] * 8 + [_f, _sz, _f], _i),
    'scp_corr_match_backward': ([_f] * 5 + [_fl, _i, _i, _i, _i, _i] + [_f] * 13 + [_f, _sz, _f], _i),
}



class VitBlock(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('ln1_w', 'ln1_b', 'qkv_w', 'qkv_b', 'proj_w', 'proj_b', 'ln2_w', 'ln2_b',
                                               'fc1_w', 'fc1_b', 'fc2_w', 'fc2_b')]


class VitWeights(ctypes.Structure):
    _fields_ = [('patch_w', ctypes.c_void_p), ('patch_b', ctypes.c_void_p), ('cls_pos0', ctypes.c_void_p),
                ('pos', ctypes.c_void_p), ('blocks', VitBlock * 12)]


_SIGNATURES.update({
    'scp_gemm_bf16_tn': ([_f, _f, _f, _f, _i, _i, _i, _f], _i),
    'scp_attention_bf16': ([_f, _f, _f, _f, _i, _i, _f], _i),
    'scp_attention_tc5': ([_f, _f, _f, _f, _i, _i, _f], _i),
    'scp_gemm_bf16x3_tn': ([_f, _f, _f, _f, _i, _i, _i, _f], _i),
    'scp_attention_x3': ([_f, _f, _f, _i, _i, _f], _i),
    'scp_vit_workspace_bytes': ([_i, _i, _i, _i], _sz),
    'scp_vit_s8_keys': ([ctypes.POINTER(VitWeights), _f, _f, _f, _i, _i, _i, _i, _i, _f, _sz, _f], _i),
    'scp_dino_argmatch': ([_f, _f, _f, _f, _i, _i, _i, _i, _f, _f], _i),
})

_pp = ctypes.POINTER(ctypes.c_void_p)
_pll = ctypes.POINTER(ctypes.c_longlong)
_SIGNATURES.update({
    'scp_image_losses_workspace_bytes': ([_i], _sz),
    'scp_image_losses_forward': ([_pp, _pll, _f, _i, _i, _i, _i, _i, _i, _f, _f, _f], _i),
    'scp_image_losses_backward': ([_pp, _pll, _f, _i, _i, _i, _i, _i, _i, _f, _f, _pp, _pll, _f, _f], _i),
})

_SIGNATURES.update({
    'scp_spmm3': ([_f] * 5 + [_i, _i, _f], _i),
    'scp_project_faces_forward': ([_f] * 6 + [_i, _i, _i, _fl] + [_f] * 4, _i),
    'scp_project_faces_backward': ([_f] * 7 + [_i, _i, _i] + [_f] * 7, _i),
})

_SIGNATURES.update({
    'scp_cycle_rows_forward': ([_f] * 8 + [_fl, _i, _i, _i, _i, _i] + [_f] * 3, _i),
    'scp_cycle_rows_backward': ([_f] * 8 + [_fl, _i, _i, _i, _i, _i] + [_f] * 4, _i),
})

_SIGNATURES.update({
    'scp_symmetry_nn_forward': ([_f] * 5 + [_i, _i, _i, _i] + [_f, _f, _f], _i),
    'scp_symmetry_nn_backward': ([_f] * 7 + [_i, _i, _i, _i] + [_f, _f], _i),
})

_SIGNATURES.update({
    'scp_color_jitter_workspace_bytes': ([_i], _sz),
    'scp_color_jitter_normalize_dparams': ([_f, _f, _i, _i, _f, _i, _f, _sz, _f], _i),
    'scp_color_jitter_normalize': ([_f, _f, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_fl), _fl, ctypes.POINTER(_fl),
                                    ctypes.POINTER(_fl), _i, _f, _sz, _f], _i),
})

_SIGNATURES.update({
    'scp_peer_buffer_bytes': ([], _sz),
    'scp_peer_buffer_create': ([_pp, ctypes.c_char_p], _i),
    'scp_peer_buffer_open': ([ctypes.c_char_p, _pp], _i),
    'scp_peer_buffer_close': ([_f, _i], _i),
    'scp_peer_exchange': ([_pp, _f, _i, _i, _i, _f, _f, _i, _f], _i),
})

_SIGNATURES.update({
    'scp_data_bbox_crop': ([_f, _f, _f, _i, _i, _i, _i, _i] + [_f] * 6 + [_f], _i),
    'scp_data_resized_crop': ([_f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _f, _f, _f, _f], _i),
})

_SIGNATURES.update({
    'scp_nhwc_maxpool3x3s2_forward': ([_f, _f, _f, _i, _i, _i, _i, _f], _i),
    'scp_nhwc_maxpool3x3s2_backward': ([_f, _f, _f, _i, _i, _i, _i, _f], _i),
    'scp_nhwc_upsample_bilinear_forward': ([_f, _f, _i, _i, _i, _i, _i, _i, _f], _i),
    'scp_nhwc_upsample2x_bilinear_backward': ([_f, _f, _i, _i, _i, _i, _f], _i),
    'scp_nhwc_l2norm_forward': ([_f, _f, _f, _i, _i, _i, _fl, _f], _i),
    'scp_nhwc_l2norm_backward': ([_f, _f, _f, _f, _i, _i, _i, _f], _i),
})

_SIGNATURES.update({
    'scp_posefit_chunks': ([_i], _sz),
    'scp_posefit_residual_table': ([_f] * 5 + [_i, _i, _i, _f, _f], _i),
    'scp_posefit_inlier_moments': ([_f] * 7 + [_i, _i, _f, _f], _i),
})

_lib = None


class ScpNativeError(RuntimeError):
    pass


def lib():
    """Loads libscp_b200.so once; raises if it has not been built (python -m self_corr_pose_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ScpNativeError(
                'native library %s is missing: run `python -m self_corr_pose_b200.build` '
                '(there is no CPU / PyTorch fallback for the hot path)' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        if handle.scp_abi_version() != ABI_VERSION:
            raise ScpNativeError('libscp_b200.so ABI %d != expected %d: rebuild' %
                                 (handle.scp_abi_version(), ABI_VERSION))
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().scp_last_error().decode('utf-8', 'replace')
        raise ScpNativeError('%s failed (code %d): %s' % (what, rc, msg))


def ptr(t):
    """Device pointer of a CUDA fp32/int tensor that must already be contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ScpNativeError('expected a CUDA tensor (the hot path has no CPU implementation)')
    if not t.is_contiguous():
        raise ScpNativeError('expected a contiguous tensor')
    return t.data_ptr()


def expect_numel(op, **buffers):
    """Raises unless every named (tensor, count) pair holds exactly the element count the kernel will index: the C-ABI
    takes bare pointers, so a short buffer would be read or written out of bounds without any diagnostic."""
    for name, (t, n) in buffers.items():
        if t is not None and t.numel() != n:
            raise ValueError('%s: `%s` has %d elements (shape %s), the kernel indexes %d'
                             % (op, name, t.numel(), tuple(t.shape), n))


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream
