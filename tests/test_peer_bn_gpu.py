"""2-GPU test of the NVLink peer-memory SyncBatchNorm (csrc/scp_peer.cu, ops/peer_sync_bn.py) against torch.nn.SyncBatchNorm
over NCCL: same outputs, input / weight gradients and running statistics on both ranks, eagerly (more exchanges than buffer
slots) and through CUDA-graph replays.  Needs two GPUs on one node (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import copy
    import torch.distributed as dist
    import torch.nn as nn
    from self_corr_pose_b200.ops import peer_sync_bn
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    ref = nn.Sequential(nn.Conv2d(3, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(), nn.Conv2d(64, 512, 3, padding=1),
                        nn.BatchNorm2d(512)).to(dev)
    ref = nn.SyncBatchNorm.convert_sync_batchnorm(ref)
    mine = copy.deepcopy(ref)
    assert peer_sync_bn.enable(mine, dev)
    assert sum(isinstance(m, peer_sync_bn.PeerSyncBatchNorm) for m in mine.modules()) == 2
    worst = 0.0
    for it in range(6):                       # 6 iterations x 2 layers x (fwd + bwd) = 24 exchanges > 4 slots
        g = torch.Generator().manual_seed(100 * it + rank)
        x = torch.randn(4 + rank, 3, 16, 16, generator=g).to(dev)          # different batch sizes per rank: counts differ
        res = []
        for net in (ref, mine):
            xi = x.clone().requires_grad_(True)
            net.zero_grad()
            y = net(xi)
            (y * y).mean().backward()
            res.append((y.detach(), xi.grad, net[1].weight.grad, net[4].running_mean.clone(), net[4].running_var.clone()))
        for a, b in zip(*res):
            worst = max(worst, float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)))
        del y, xi, res      # no live autograd graph (its AccumulateGrad nodes are tied to the default stream) into the capture
    # CUDA graph (as Trainer.capture does it: gradients pre-allocated and zeroed inside the graph, warm-up on a side stream):
    # capture one forward + backward, replay three times with new inputs
    x_static = torch.randn(4, 3, 16, 16, device=dev)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            mine.zero_grad(set_to_none=False)
            (mine(x_static) ** 2).mean().backward()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, capture_error_mode='thread_local'):      # NCCL's watchdog thread polls events meanwhile
        for p in mine.parameters():
            p.grad.zero_()
        y_static = mine(x_static)
        (y_static ** 2).mean().backward()
    for it in range(3):
        g = torch.Generator().manual_seed(900 + 10 * it + rank)
        xn = torch.randn(4, 3, 16, 16, generator=g).to(dev)
        ref.load_state_dict(mine.state_dict())          # same running statistics going in
        x_static.copy_(xn)
        graph.replay()
        ref.zero_grad()
        yr = ref(xn)
        (yr ** 2).mean().backward()
        torch.cuda.synchronize()
        worst = max(worst, float((y_static - yr).abs().max() / yr.abs().max()),
                    float((mine[0].weight.grad - ref[0].weight.grad).abs().max() / ref[0].weight.grad.abs().max()))
    out[rank] = worst
    dist.barrier()
    dist.destroy_process_group()


def test_peer_sync_batchnorm_matches_nccl_sync_batchnorm():
    os.environ.setdefault('SCP_SYNTHETIC_WEIGHTS', '1')
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    print('PARITY peer SyncBatchNorm vs NCCL SyncBatchNorm: worst relative difference per rank', dict(out))
    assert len(out) == world and max(out.values()) < 1e-5
