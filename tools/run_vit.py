import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from self_corr_pose_b200.model.module.network.dino import DINO
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = DINO().cuda()
img = torch.rand(B, 3, 256, 256, device='cuda')
for _ in range(2):
    f = net(img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); f = net(img); e1.record(); torch.cuda.synchronize()
print('vit ms', e0.elapsed_time(e1))
