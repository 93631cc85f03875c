"""ViT-S/8 layer-9 key extractor and its attention kernel alone, CUDA events (B = 64, 256x256)."""
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


net = DINO().cuda()
img = torch.rand(B, 3, 256, 256, device='cuda')
T = 1025
q = torch.randn(B * 6, T, 64, device='cuda').to(torch.bfloat16)
Tp = (T + 7) // 8 * 8
vt = torch.zeros(B * 6, 64, Tp, device='cuda', dtype=torch.bfloat16)
vt[:, :, :T] = q.transpose(1, 2)
o = torch.empty(B, T, 384, device='cuda', dtype=torch.bfloat16)
st = _lib.stream_ptr(img.device)
res = {'variant': os.environ.get('SCP_VIT_ATTENTION', 'default(2)'), 'B': B}
res['vit_ms'] = timeit(lambda: net(img), 5)
res['attention_ms'] = timeit(lambda: L.scp_attention_tc5(_lib.ptr(q), _lib.ptr(q), _lib.ptr(vt), _lib.ptr(o), B, T, st))
res['attention_tflops'] = 4.0 * T * T * 64 * 6 * B / res['attention_ms'] / 1e9
res['vit_tflops'] = 47.62e9 * B / res['vit_ms'] / 1e9
print(json.dumps(res))
