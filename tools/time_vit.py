"""ViT-S/8 layer-9 key extractor and its attention kernel alone, CUDA events (B = 64, 256x256)."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from self_corr_pose_b200 import _lib
from self_corr_pose_b200.model.module.network.dino import DINO

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = _lib.lib()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


img = torch.rand(B, 3, 256, 256, device='cuda')
T = 1025
q = torch.randn(B * 6, T, 64, device='cuda').to(torch.bfloat16)
Tp = (T + 7) // 8 * 8
vt = torch.zeros(B * 6, 64, Tp, device='cuda', dtype=torch.bfloat16)
vt[:, :, :T] = q.transpose(1, 2)
o = torch.empty(B, T, 384, device='cuda', dtype=torch.bfloat16)
st = _lib.stream_ptr(img.device)
res = {'variant': os.environ.get('SCP_VIT_ATTENTION', 'default(2)'), 'B': B}
for prec in ('x3', 'bf16'):
    net = DINO(precision=prec).cuda()
    res['vit_%s_ms' % prec] = timeit(lambda: net(img), 5)
    res['vit_%s_tflops_algorithmic' % prec] = 47.62e9 * B / res['vit_%s_ms' % prec] / 1e9
from self_corr_pose_b200.model.module.network.dino import split_bf16_i32
qk = torch.randn(B * T, 1536, device='cuda').to(torch.bfloat16)
vt2 = torch.zeros(2, B * 384, Tp, device='cuda', dtype=torch.bfloat16)
vt2[:, :, :T] = torch.randn(2, B * 384, T, device='cuda').to(torch.bfloat16)
o3 = torch.empty(B * T, 768, device='cuda', dtype=torch.bfloat16)
res['attention_x3_ms'] = timeit(lambda: L.scp_attention_x3(_lib.ptr(qk), _lib.ptr(vt2), _lib.ptr(o3), B, T, st))
M = B * T
A3 = torch.randn(M, 768, device='cuda').to(torch.bfloat16)
W3 = torch.randn(1152, 768, device='cuda').to(torch.bfloat16)
C3 = torch.empty(M, 1152, device='cuda')
res['gemm_qkv_x3_ms'] = timeit(lambda: L.scp_gemm_bf16x3_tn(_lib.ptr(A3), _lib.ptr(W3), None, _lib.ptr(C3), M, 1152, 384, st))
res['gemm_qkv_bf16_ms'] = timeit(lambda: L.scp_gemm_bf16_tn(_lib.ptr(A3), _lib.ptr(W3), None, _lib.ptr(C3), M, 1152, 768, st))
res['attention_ms'] = timeit(lambda: L.scp_attention_tc5(_lib.ptr(q), _lib.ptr(q), _lib.ptr(vt), _lib.ptr(o), B, T, st))
res['attention_tflops'] = 4.0 * T * T * 64 * 6 * B / res['attention_ms'] / 1e9
print(json.dumps(res))
