#!/usr/bin/env python
"""bench.py -- images/sec of the self-corr-pose training step (forward + backward) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B_per_gpu] [--impl b200|reference]
                    [--workload trainer|hotpath|config1]

Default workload: `Trainer(opts).step(batch)` -- the reference's training-loop body (model/trainer.py:118-125 around
MeshNet.forward, model/model.py:61-152): ResNet-18 encoder (both passes), fused correspondence, texture sampling, 4 SoftRas
renders, silhouette / texture / depth / match / imatch losses, DINO ViT-S/8 features (fp32-class precision) + pseudo-matches
+ pre-training cycle loss, rotation-cycle loss, symmetry / Laplacian / deformation regularisers, backward, ONE flat-buffer
gradient all-reduce (N > 1), clipping, AdamW + OneCycle -- over a synthetic batch of 256x256 images, 64 per GPU, the
1280-vertex / 2556-face category mesh, 64x64 correspondence map, C = 64 (BASELINE.json configs[2]/[3]).
`--workload hotpath` times the round-1 unit (model.py:73-134 with the encoder outputs as inputs, one CUDA graph); it is
also reported inside the default line as `hotpath`.  `--workload config1` = BASELINE configs[1] (B = 32: DINO ViT-S/8
features + P=1024/4096 x N=1280 correspondence + pre-training cycle block).  Prints ONE JSON line.

--impl reference: the same workload in the reference's formulation on the host cores (oracle/trainer_cpu.py /
oracle/hotpath_cpu.py: the reference has no CPU path of its own), on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('SCP_SYNTHETIC_WEIGHTS', '1')   # no checkpoints offline: synthetic weights, stated in `data`
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=64, help='images per GPU (batch_size x repeat, repeat = 4)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--mesh', default='uv1280')
    ap.add_argument('--workload', default='trainer', choices=['trainer', 'hotpath', 'config1'])
    ap.add_argument('--vit-precision', default='x3', choices=['x3', 'bf16'],
                    help="x3 = fp32-class split-bf16 products (parity mode, default); bf16 = labelled fast mode")
    ap.add_argument('--cpu-batch', type=int, default=2, help='images in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-breakdown', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step eagerly instead of replaying a CUDA graph')
    ap.add_argument('--no-overlap', action='store_true', help='run the DINO ViT on the main stream (no side-stream overlap)')
    return ap.parse_args()


def load_mesh(name):
    from self_corr_pose_b200 import synthetic
    if name == 'uv1280':
        return synthetic.uv_sphere()
    if name.startswith('ico'):
        return synthetic.icosphere(int(name[3:]))
    return synthetic.load_prior(name)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def mesh_prior_path(name):
    return 'synthetic:uv1280' if name == 'uv1280' else 'config/%s_wild6d/%s.obj' % (name, name)


WORKLOAD_TEXT = {
    'trainer': 'Trainer.step: MeshNet.forward (ResNet-18 encoder x2 passes, fused correspondence P=4096 x N=%d x C=64, texture, '
               '4 SoftRas renders nf=%d, mask/texture/depth/match/imatch losses, DINO ViT-S/8 layer-9 keys + pseudo-matches + '
               'pre-train cycle loss, rotation-cycle loss, symmetry/Laplacian/deform regularisers) + backward + flat '
               'gradient all-reduce + clip + AdamW/OneCycle; laptop_wild6d flag values',
    'hotpath': 'configs[2] hot path fwd+bwd = fused correspondence (P=4096,N=%d,C=64) -> texture -> 4 SoftRas renders (nf=%d) '
               '-> mask/texture/depth/match/imatch losses -> DINO ViT-S/8 layer-9 keys + pseudo-matches + pre-train cycle '
               'loss; encoder outputs are inputs',
    'config1': 'configs[1]: DINO ViT-S/8 layer-9 keys of the batch + correspondence match fwd+bwd at P=1024 and P=4096 '
               '(N=%d, C=64) + pre-training cycle block (arg-match, top-k, cycle rows fwd+bwd); nf=%d unused',
}


def cpu_step_rate(args, opts_kw, label):
    """The workload in the reference's formulation on the host cores, bounded sample; returns (cpu_baseline object, seconds)."""
    import torch
    from oracle import hotpath_cpu as H
    from self_corr_pose_b200.hotpath import default_opts
    from self_corr_pose_b200.model.module.network.vit_weights import synthetic_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bc = max(2, args.cpu_batch)
    opts = default_opts(**dict(opts_kw, batch_size=Bc // 2, repeat=2, shape_prior_path=mesh_prior_path(args.mesh)))
    v, f = load_mesh(args.mesh)
    sd = synthetic_state_dict(0)
    use_ref = os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'libsoftras_ref_cpu.so'))
    if args.workload == 'trainer':
        from types import SimpleNamespace
        from oracle import trainer_cpu as TC
        from self_corr_pose_b200 import synthetic
        from self_corr_pose_b200.model.module.renderer import Renderer
        from self_corr_pose_b200.model.trainer import Trainer
        cpu = TC.CpuTrainer(opts, sd, use_ref=use_ref)
        mesh = SimpleNamespace(mean_v=torch.from_numpy(v), faces=torch.from_numpy(f), texture_type='vertex')
        with H.cpu_rasterizer():
            batch = synthetic.make_trainer_batch(opts, v, f, Bc, device='cpu', seed=0, renderer=Renderer(opts, mesh))
        shaper = Trainer(opts)
        shaper.device = torch.device('cpu')
        data = shaper.batch_reshape(batch)
        t0 = time.time()
        cpu.step(data)
        dt = time.time() - t0
        what = 'Trainer.step (encoder x2 in torch CPU ops, symmetry / rotation-cycle terms, clip, AdamW)'
    else:
        data, enc = H.make_batch_cpu(opts, v, f, Bc, seed=0)
        t0 = time.time()
        H.step(opts, torch.from_numpy(v), torch.from_numpy(f), data, enc, sd, use_ref=use_ref, all_vit_blocks=False)
        dt = time.time() - t0
        what = 'hot path'
    return {'value': Bc / dt, 'unit': 'images/sec', 'cores': cores, 'kind': 'reference' if use_ref else 'port',
            'sample': '%s: 1 %s step of %d images 256x256 (%s mesh), reference formulation on CPU: torch ops + '
                      '%s SoftRas over all faces per pixel, DINO ViT on the 4x duplicated pair batch up to layer 9; %.1f s'
                      % (label, what, Bc, args.mesh, 'the reference kernel source built for the host (oracle/_ref)' if use_ref
                         else 'C restatement (oracle/softras_oracle.c)', dt)}, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if args.workload == 'config1':
        args.workload = 'hotpath'
    opts_kw = dict(img_size=256, corr_h=64, corr_w=64)
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        cb, dt = cpu_step_rate(args, opts_kw, 'impl=reference')
        if i >= args.warmup:
            vals.append(cb['value'])
        last = cb
        if i == 0 and dt * (args.warmup + args.steps) > 240:   # keep the whole run within a few minutes
            vals = [cb['value']]
            break
    v = sum(vals) / len(vals)
    last['value'] = v
    vv, ff = load_mesh(args.mesh)
    print(json.dumps({
        'impl': 'reference', 'metric': 'images/sec fwd+bwd (feat+corr+render+loss) 256x256', 'value': v,
        'unit': 'images/sec', 'n_gpus': args.gpus, 'steps': len(vals), 'warmup': args.warmup,
        'ms_per_step': 1e3 * max(2, args.cpu_batch) / v, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_TEXT[args.workload] % (vv.shape[0], ff.shape[0]),
                   'sample': 'CPU sample of %d images' % max(2, args.cpu_batch), 'mesh': args.mesh, 'img_size': 256},
        'cpu_baseline': last,
        'e2e': {'value': v, 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def kernel_breakdown(torch, hot, data, enc, B, peaks):
    """CUDA-event times of the individual kernels (each launched alone on the current stream) and their
    rooflines; returns (dominant roofline object, list)."""
    from self_corr_pose_b200 import _lib
    from self_corr_pose_b200.soft_renderer import functional as srf
    from self_corr_pose_b200.model.util.loss_utils import project_to_screen
    dev = data[0].device
    L = _lib.lib()
    hbm, tf = peaks.get('hbm_gbs', 6650.0), peaks.get('bf16_tflops', 1590.0)
    which = 'measured' if peaks else 'fallback'

    def timeit(fn, n=7):
        """Device time of one call of `fn`: the call is captured into a CUDA graph (no host launch overhead, same as the
        graph-replayed step) and replayed n times between CUDA events; median of 3 such measurements."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = None
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                keep = fn()   # noqa: F841 -- keeps the captured outputs alive
            run = graph.replay
        except Exception:   # noqa: BLE001
            graph, run = None, fn
        ts = []
        for _ in range(3):
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / n)
        del graph
        return sorted(ts)[1]

    out = []
    img, mask, depth, foc, pp = data
    img_feat, mesh_feat, pred_v, rot, trans = [t.detach() for t in enc]
    N = pred_v.shape[1]
    nf = hot.mesh.faces.shape[0]
    is_ = hot.opts.img_size
    # --- SoftRas: soft-texture render (sigma 1e-3), forward and backward kernels
    with torch.no_grad():
        sv = project_to_screen(pred_v.clone(), foc, pp, rot, trans)
        sv = torch.stack((sv[..., 0], sv[..., 1], sv[..., 2] + 2.7320508), -1)
        fv = srf.face_vertices(sv, hot.mesh.faces[None].repeat(B, 1, 1)).contiguous()
        tex = srf.face_vertices(torch.rand_like(sv), hot.mesh.faces[None].repeat(B, 1, 1)).contiguous()
    bytes_f = B * (72 * nf + 24 * is_ * is_)
    bytes_b = B * (144 * nf + 40 * is_ * is_)
    def candidate_pairs(sigma):
        """SURVEY.md 8(d) secondary figure: (pixel, face) pairs inside the faces' inflated bounding boxes, summed over the
        batch -- what a traversal has to evaluate (dense would be B * nf * is^2).  None if anything goes wrong."""
        try:
            m = (9.2102404 * sigma) ** 0.5                       # sqrt(ln(1/dist_eps - 1) * sigma), dist_eps = 1e-4
            lo, hi = fv[..., :2].amin(2) - m, fv[..., :2].amax(2) + m           # B, nf, 2
            ext = (hi.clamp(-1, 1) - lo.clamp(-1, 1)).clamp_min(0) * (is_ / 2)  # pixels
            return float((ext[..., 0] * ext[..., 1]).double().sum())
        except Exception:   # noqa: BLE001
            return None

    for name in ('softras_softtex', 'softras_depth+nocs'):
        fvg = fv.clone().requires_grad_(True)
        if name == 'softras_softtex':
            kw = dict(image_size=is_, background_color=[1, 1, 1], sigma_val=1e-3, gamma_val=1e-2, aggr_func_rgb='softmax',
                      texture_type='vertex')
            run = lambda: srf.soft_rasterize(fvg, tex, **kw)
            bf = bytes_f
        else:   # depth render + NOCS map in one traversal (two output sets)
            run = lambda: srf.soft_rasterize_dual(fvg, tex, tex, image_size=is_, sigma_val=1e-4, gamma_val=1e-4)[0]
            bf = bytes_f + B * (36 * nf + 24 * is_ * is_)
        t_f = timeit(run)
        o = run()
        g = torch.randn_like(o)
        t_b = timeit(lambda: torch.autograd.grad(o, fvg, g, retain_graph=True))
        out.append(dict(kernel=name + '_fwd (pack+forward_kernel)', ms=t_f, bound='hbm', achieved=bf / t_f / 1e6,
                        peak=hbm, unit='GB/s', launches_per_step=1,
                        ncu_name='softras::forward_kernel<1, 1> #0' if 'softtex' in name else 'softras::forward_kernel<2, 1> #0'))
        out.append(dict(kernel=name.replace('+nocs', '') + '_bwd (pack+backward_face_kernel)', ms=t_b, bound='hbm',
                        achieved=bytes_b / t_b / 1e6, peak=hbm, unit='GB/s', launches_per_step=1,
                        ncu_name='softras::backward_face_kernel<1, 1> #%d' % (0 if 'softtex' in name else 1)))
        pairs = candidate_pairs(1e-3 if 'softtex' in name else 1e-4)
        if pairs:
            for k, t in ((out[-2], t_f), (out[-1], t_b)):
                k['candidate_pairs'] = pairs
                k['gpairs_per_s'] = pairs / t / 1e6
    # --- correspondence
    from self_corr_pose_b200.ops.corr_match import corr_match
    import torch.nn.functional as F
    hf, wf = hot.opts.corr_h, hot.opts.corr_w
    P, C = hf * wf, hot.opts.n_corr_feat
    md = F.interpolate(mask[:, None], (hf, wf), mode='nearest').reshape(B, -1)
    a = img_feat.clone().requires_grad_(True)
    m = mesh_feat.clone().requires_grad_(True)
    fwd = lambda: corr_match(a, m, md, pred_v, hot.corr_net.meshgrid, 10.0, hf, wf, want_full=False, want_pool=True)
    t_f = timeit(fwd)
    _, pool, mt, im, _A = fwd()
    gs = [torch.randn_like(pool), torch.randn_like(mt), torch.randn_like(im)]
    t_b = timeit(lambda: torch.autograd.grad([pool, mt, im], [a, m], gs, retain_graph=True))
    bytes_f = B * (4 * (C * P + N * C + P + 3 * N) + 4 * (P * N // 4 + 2 * N + 3 * P))
    bytes_b = B * 4 * (2 * C * P + 2 * N * C + P * N // 4 + 6 * P + 7 * N)
    out.append(dict(kernel='correspondence forward: gemm_rowstats_kernel x2 (tcgen05 kind::tf32, S and S^T) + prep / fill / combine', ms=t_f, bound='hbm', achieved=bytes_f / t_f / 1e6, peak=hbm,
                    unit='GB/s', launches_per_step=1, ncu_name='gemm_rs::gemm_rowstats_kernel<corr_tc::EpiCols<1>> #0'))
    out.append(dict(kernel='corr_bwd_rows_kernel<fused cols> (+blocklist)', ms=t_b, bound='hbm', achieved=bytes_b / t_b / 1e6,
                    peak=hbm, unit='GB/s', launches_per_step=1, ncu_name='corr::corr_bwd_rows_kernel<1> #0'))
    # --- ViT: whole extractor (a chain of 68 launches), its attention kernel and the QKV GEMM alone, in the precision the
    # step runs (x3: three tensor-core products per logical product -> peak = measured bf16 / 3, labelled derived)
    net = hot.pretrain_corr_net.net
    x3 = net.precision == 'x3'
    tf_eff = tf / 3 if x3 else tf
    tf_note = 'derived: measured bf16 / 3 (x3 split products)' if x3 else which
    t_v = timeit(lambda: net(img), n=5)
    out.append(dict(kernel='vit_s8_keys[%s] (chain of 68 launches)' % net.precision, ms=t_v, bound='tensor',
                    achieved=47.62e9 * B / t_v / 1e9, peak=tf_eff, unit='TFLOP/s', launches_per_step=1, chain=True,
                    peak_source=tf_note))
    T = (is_ // 8) ** 2 + 1
    Tp = (T + 7) // 8 * 8
    st = _lib.stream_ptr(dev)
    M = B * T
    if x3:
        from self_corr_pose_b200.model.module.network.dino import split_bf16_i32
        qk = split_bf16_i32(torch.randn(M, 768, device=dev))            # proper (hi, lo) pairs of N(0,1) queries / keys
        vt = torch.zeros(2, B * 384, Tp, device=dev, dtype=torch.bfloat16)
        vv = torch.randn(B * 384, T, device=dev)
        vt[0, :, :T] = vv.to(torch.bfloat16)
        vt[1, :, :T] = (vv - vt[0, :, :T].float()).to(torch.bfloat16)
        o = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
        t_a = timeit(lambda: L.scp_attention_x3(_lib.ptr(qk), _lib.ptr(vt), _lib.ptr(o), B, T, st))
        out.append(dict(kernel='fa3_fwd_kernel (tcgen05 flash attention, split operands)', ms=t_a, bound='tensor',
                        achieved=4.0 * T * T * 64 * 6 * B / t_a / 1e9, peak=tf_eff, unit='TFLOP/s', launches_per_step=9,
                        ncu_name='fa3::fa3_fwd_kernel<4, 0> #0', peak_source=tf_note))
        A = split_bf16_i32(torch.randn(M, 384, device=dev))
        W = split_bf16_i32(torch.randn(1152, 384, device=dev))
        Cc = torch.empty(M, 1152, device=dev)
        t_g = timeit(lambda: L.scp_gemm_bf16x3_tn(_lib.ptr(A), _lib.ptr(W), None, _lib.ptr(Cc), M, 1152, 384, st))
        out.append(dict(kernel='gemm_bf16_tn_kernel<.,3> (qkv shape, fp32 out)', ms=t_g, bound='tensor',
                        achieved=2.0 * M * 1152 * 384 / t_g / 1e9, peak=tf_eff, unit='TFLOP/s', launches_per_step=9,
                        peak_source=tf_note))
    else:
        q = torch.randn(B * 6, T, 64, device=dev).to(torch.bfloat16)
        vt = torch.zeros(B * 6, 64, Tp, device=dev, dtype=torch.bfloat16)
        vt[:, :, :T] = q.transpose(1, 2)
        o = torch.empty(B, T, 384, device=dev, dtype=torch.bfloat16)
        t_a = timeit(lambda: L.scp_attention_tc5(_lib.ptr(q), _lib.ptr(q), _lib.ptr(vt), _lib.ptr(o), B, T, st))
        out.append(dict(kernel='fa2_fwd_kernel (tcgen05 flash attention)', ms=t_a, bound='tensor',
                        achieved=4.0 * T * T * 64 * 6 * B / t_a / 1e9, peak=tf, unit='TFLOP/s', launches_per_step=9,
                        ncu_name='fa2::fa2_fwd_kernel #0'))
        A = torch.randn(M, 384, device=dev).to(torch.bfloat16)
        W = torch.randn(1152, 384, device=dev).to(torch.bfloat16)
        Cc = torch.empty(M, 1152, device=dev)
        t_g = timeit(lambda: L.scp_gemm_bf16_tn(_lib.ptr(A), _lib.ptr(W), None, _lib.ptr(Cc), M, 1152, 384, st))
        out.append(dict(kernel='gemm_bf16_tn_kernel (qkv shape, fp32 out)', ms=t_g, bound='tensor',
                        achieved=2.0 * M * 1152 * 384 / t_g / 1e9, peak=tf, unit='TFLOP/s', launches_per_step=9))
    for k in out:
        k['frac'] = k['achieved'] / k['peak']
        k['step_ms'] = k['ms'] * k['launches_per_step']
    dom = max((k for k in out if not k.get('chain')), key=lambda k: k['step_ms'])    # largest single kernel per step
    # DRAM bytes of one launch of that kernel -- and, because these kernels are issue / tensor bound rather than HBM
    # bound, its issue-slot and tensor-pipe utilisation -- from the committed `ncu --set full` capture (profiles/)
    traffic, ncu = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        for k in out:
            rec = tr.get(k.get('ncu_name', ''))
            if rec:
                k['ncu'] = {m: rec[m] for m in ('dram_bytes', 'issue_active_pct', 'tensor_pipe_pct', 'dram_throughput_pct')
                            if m in rec}
        rec = tr.get(dom.get('ncu_name', ''), {})
        traffic = rec.get('dram_bytes')
        ncu = {m: rec[m] for m in ('issue_active_pct', 'tensor_pipe_pct', 'dram_throughput_pct') if m in rec} or None
    except Exception:
        pass
    roof = {'kernel': dom['kernel'], 'bound': dom['bound'], 'achieved': dom['achieved'], 'peak': dom['peak'],
            'unit': dom['unit'], 'frac': dom['frac'], 'traffic': traffic, 'ncu': ncu, 'peak_source': dom.get('peak_source', which),
            'ms_per_launch': dom['ms'], 'launches_per_step': dom['launches_per_step']}
    return roof, out


class Timer:
    """K calls of fn between CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, torch, dist, dev, world):
        self.torch, self.dist, self.dev, self.world = torch, dist, dev, world

    def __call__(self, fn, k):
        torch = self.torch
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)


def hotpath_arm(args, torch, dist, dev, world, rank, timed):
    """Round-1 unit: HotPath step as ONE CUDA graph.  Returns dict(value, ms, e2e..., objects for the kernel breakdown)."""
    from self_corr_pose_b200.hotpath import HotPath, default_opts
    from self_corr_pose_b200.model.module.renderer import Renderer
    from self_corr_pose_b200 import synthetic
    B = args.batch
    opts = default_opts(img_size=256, corr_h=64, corr_w=64, batch_size=B // 4, repeat=4)
    v, f = load_mesh(args.mesh)
    mean_v, faces = torch.from_numpy(v), torch.from_numpy(f)
    hot = HotPath(opts, mean_v, faces, device=dev, overlap_vit=not args.no_overlap)
    data, enc = synthetic.make_batch(opts, v, f, B, device=dev, seed=rank, renderer=Renderer(opts, hot.mesh))
    graphed = None
    if not args.no_graph:
        try:   # whole step (forward + backward, ~900 launches) as ONE CUDA graph over static buffers
            graphed = hot.capture(data, enc)
        except Exception as e:   # noqa: BLE001 -- report and continue eagerly
            print('CUDA graph capture failed, running eagerly: %r' % (e,), file=sys.stderr)

    def step(d):
        if graphed is not None:
            if d is not data:
                graphed.load(data=d)
            return graphed.replay()
        return hot.step(d, enc)[0]

    for _ in range(max(3, args.warmup)):
        loss = step(data)
    ms = timed(lambda: step(data), args.steps)

    # end to end through the public call with HOST buffers: H2D of the batch + step + D2H of the loss
    host = [t.detach().cpu().pin_memory() for t in data]
    h2d = sum(t.numel() * t.element_size() for t in host)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(dev)
    staging = [tuple(torch.empty_like(t) for t in data) for _ in range(2)]
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {'i': 0}

    def upload(slot):
        copy_stream.wait_event(consumed[slot])          # the step that read this slot has copied it out
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(staging[slot], host):
                dst.copy_(src, non_blocking=True)
            uploaded[slot].record(copy_stream)

    for ev in consumed:
        ev.record(torch.cuda.current_stream(dev))
    upload(0)

    def e2e_step():
        slot = state['i'] & 1
        state['i'] += 1
        main = torch.cuda.current_stream(dev)
        main.wait_event(uploaded[slot])
        if graphed is not None:
            graphed.load(data=staging[slot])            # device-to-device into the graph's static buffers
            consumed[slot].record(main)
            upload(slot ^ 1)
            total = step(data)
        else:
            upload(slot ^ 1)
            total = step(staging[slot])
            consumed[slot].record(main)
        loss_host.copy_(total.detach(), non_blocking=True)
        main.synchronize()
        return float(loss_host)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    return dict(ms=ms / args.steps, ms_e2e=ms_e2e / args.steps, h2d=h2d, d2h=4, loss=float(loss.detach()), hot=hot, data=data,
                enc=enc, graph=graphed is not None, launches=hot.GPU_LAUNCHES, N=v.shape[0], nf=f.shape[0])


def trainer_arm(args, torch, dist, dev, world, rank, local, timed):
    """Trainer.step on a device-resident batch (value) and on a pinned host batch with the loss read back (e2e)."""
    from self_corr_pose_b200 import synthetic
    from self_corr_pose_b200.hotpath import default_opts
    from self_corr_pose_b200.model.trainer import Trainer
    from self_corr_pose_b200.model.module.renderer import Renderer
    B = args.batch
    torch.backends.cudnn.benchmark = True               # train.py:21 of the reference
    opts = default_opts(img_size=256, corr_h=64, corr_w=64, batch_size=B // 4, repeat=4, local_rank=local, ngpu=world,
                        shape_prior_path=mesh_prior_path(args.mesh))
    torch.manual_seed(0)
    tr = Trainer(opts)
    model = tr.define_model()
    model.overlap_rotation = os.environ.get('SCP_OVERLAP_ROTATION', '1') != '0'     # A/B switch of the second side stream
    model.overlap_vit = not args.no_overlap
    v, f = load_mesh(args.mesh)
    batch = synthetic.make_trainer_batch(opts, v, f, B, device=dev, seed=rank, renderer=Renderer(opts, model.mesh))
    graphed = False
    if not args.no_graph:
        try:        # zero-grad + forward + backward of the step as ONE CUDA graph (Trainer.capture)
            tr.capture(batch, warmup=max(3, args.warmup))
            graphed = True
        except Exception as e:   # noqa: BLE001 -- report and continue eagerly
            print('CUDA graph capture of Trainer.step failed, running eagerly: %r' % (e,), file=sys.stderr)
    step_fn = tr.step_graphed if graphed else tr.step
    for _ in range(max(3, args.warmup)):
        total, aux, _ = step_fn(batch)
    prof = os.environ.get('SCP_BENCH_CUDA_PROFILER') == '1'     # ncu --profile-from-start off: launch list of the timed steps only
    if prof:
        torch.cuda.profiler.start()
    ms = timed(lambda: step_fn(batch), args.steps)
    if prof:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    host = {k: (t.detach().cpu().pin_memory() if k not in ('center', 'length') else t) for k, t in batch.items()}
    h2d = sum(t.numel() * t.element_size() for k, t in host.items() if k not in ('center', 'length'))
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    state = {'slot': tr.stage(host) if graphed else None}

    def e2e_step():     # public call with HOST buffers: every step uploads one full batch from pinned memory and reads the
        if graphed:     # loss back; the upload of the NEXT step's batch (copy stream) overlaps this step's compute
            slot = state['slot']
            state['slot'] = tr.stage(host)
            total, _, _ = tr.step_graphed(slot)
        else:
            total, _, _ = tr.step(host)         # batch_reshape uploads the batch
        loss_host.copy_(total.detach(), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return float(loss_host)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    comm = None
    if world > 1:       # the step's one data-path collective, timed alone on the real flat gradient buffer
        red = tr.reducer
        nbytes = red.flat.numel() * 4
        t = timed(lambda: dist.all_reduce(red.flat), 10) / 10
        n_sync_bn = sum(1 for m in model.modules() if isinstance(m, torch.nn.SyncBatchNorm))
        comm = {'allreduce_bytes': nbytes, 'allreduce_ms': t, 'bus_gbs': 2 * (world - 1) / world * nbytes / t / 1e6,
                'bus_gbs_reference_point': 725.0, 'frac_of_step': t / (ms / args.steps),
                'sync_batchnorm_layers': n_sync_bn,
                'sync_batchnorm_transport': 'NVLink peer-memory exchange kernel (csrc/scp_peer.cu)' if tr.peer_bn else 'NCCL',
                'sync_batchnorm_collectives_per_step': n_sync_bn * 2 * 2,   # (stats gather fwd + reduce bwd) x 2 encoder passes
                'note': 'one all_reduce(SUM) of the flat fp32 gradient buffer per step (dist.FlatGradReducer), not overlapped '
                        'with backward; SyncBatchNorm collectives as in the reference (trainer.py:66)'}
    return dict(ms=ms / args.steps, ms_e2e=ms_e2e / args.steps, h2d=h2d, d2h=4, loss=float(total.detach()), comm=comm,
                launches=model.GPU_LAUNCHES, N=v.shape[0], nf=f.shape[0], graph=graphed,
                aux={k: float(x) for k, x in aux.items()})


def config1_arm(args, torch, dev, timed):
    """BASELINE configs[1]: B = 32 images -> ViT features; correspondence fwd+bwd at P = 1024 and 4096, N = 1280, C = 64;
    pre-training cycle block on pointcorr (B, 4096, N)."""
    import torch.nn.functional as F
    from self_corr_pose_b200 import synthetic
    from self_corr_pose_b200.hotpath import default_opts
    from self_corr_pose_b200.model.module.correspondence import Correspondence
    from self_corr_pose_b200.model.module.pretrained_corr import PretrainedCorrespondence
    from types import SimpleNamespace
    B = args.batch
    g = torch.Generator().manual_seed(0)
    v, f = load_mesh(args.mesh)
    N, C = v.shape[0], 64
    img = torch.rand(B, 3, 256, 256, generator=g).to(dev)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 256), torch.linspace(-1, 1, 256), indexing='ij')
    mask = ((xx ** 2 + yy ** 2) < 0.36).float()[None].repeat(B, 1, 1).to(dev)
    pred_v = (torch.from_numpy(v)[None] + 0.01 * torch.randn(B, N, 3, generator=g)).to(dev)
    mesh_feat = F.normalize(torch.relu(torch.randn(B, N, C, generator=g)), 2, -1).to(dev).requires_grad_(True)
    nets, feats = {}, {}
    for hf in (32, 64):
        opts = default_opts(img_size=256, corr_h=hf, corr_w=hf, batch_size=B // 4, repeat=4)
        nets[hf] = Correspondence(opts, dev)
        feats[hf] = F.normalize(torch.randn(B, C, hf * hf, generator=g), 2, 1).to(dev).requires_grad_(True)
    opts = default_opts(img_size=256, corr_h=64, corr_w=64, batch_size=B // 4, repeat=4)
    pre = PretrainedCorrespondence(opts, SimpleNamespace(), device=dev).to(dev)
    depth_weight = torch.ones(B, N, device=dev)

    def step():
        for t in (mesh_feat, feats[32], feats[64]):
            t.grad = None
        loss = 0
        for hf in (32, 64):
            pointcorr, match, imatch = nets[hf].match_lowres(feats[hf], mesh_feat, mask, pred_v)
            loss = loss + match.square().mean() + imatch.square().mean()
        cyc = pre.compute_cycle_loss(img, mask, depth_weight, pointcorr, pooled=True, A=nets[64].pool_A)
        loss = loss + cyc[0]
        loss.backward()
        return loss
    for _ in range(max(3, args.warmup)):
        loss = step()
    ms = timed(step, args.steps)
    return dict(ms=ms / args.steps, ms_e2e=None, h2d=0, d2h=0, loss=float(loss.detach()), launches=68 + 2 * 8 + 2 + 2,
                N=N, nf=f.shape[0])


def data_path_figure(torch, B, peaks, with_cpu):
    """the loader's per-frame work after the file decode (mask bounding box, crop box + intrinsics, three
    resized crops) for one batch of decoded 640 x 480 frames resident in HBM -> the 256 x 256 training batch, on the GPU, and
    the same statements (oracle/data_cpu.py = the reference's __getitem__ body) on ONE host core for a few frames."""
    import numpy as np
    from oracle import data_cpu
    from self_corr_pose_b200.ops import crop_resize
    H, W, S = 480, 640, 256
    frames, Ks = data_cpu.synthetic_frames(8, H, W, seed=0)
    rep = (B + 7) // 8
    img = torch.from_numpy(np.stack([f[0] for f in frames])).cuda().repeat(rep, 1, 1, 1)[:B].contiguous()
    mask = torch.from_numpy(np.stack([f[1] for f in frames])).cuda().repeat(rep, 1, 1)[:B].contiguous()
    depth = torch.from_numpy(np.stack([f[2] for f in frames]).view(np.int16)).cuda().repeat(rep, 1, 1)[:B].contiguous()
    K = np.stack(Ks)
    intr = torch.from_numpy(np.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]], 1)).cuda().repeat(rep, 1)[:B].contiguous()
    rs = torch.from_numpy(np.random.RandomState(1).uniform(1.2, 1.5, size=(B, 2))).cuda()
    out = (torch.empty(B, 3, S, S, device='cuda'), torch.empty(B, 1, S, S, device='cuda'), torch.empty(B, 1, S, S, device='cuda'))
    res = {}
    for aa in (False, True):
        def run():
            box = crop_resize.bbox_crop(mask, rs, intr, S)
            crop_resize.resized_crop(img, mask, depth, box['crop'], S, bgr=True, antialias=aa, out=out)
            return box
        for _ in range(3):
            box = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        res[aa] = e0.elapsed_time(e1) / n
    crop = box['crop'].cpu().numpy().astype(np.int64)
    # algorithmic bytes: the mask once (bounding box) + the crop's share of the three frames + the five fp32 output planes
    inside = (np.clip(crop[:, 2], 0, H) * np.clip(crop[:, 3], 0, W)).sum()
    alg = B * H * W + inside * (3 + 1 + 2) + B * 5 * S * S * 4
    peak = float(peaks.get('hbm_gbs', 6551.0))
    fig = {'workload': 'loader after decode: %d frames %dx%d (u8 BGR + u8 mask + u16 depth) -> bbox, crop box, intrinsics, 3 resized '
                       'crops -> (img, mask, depth) %dx%d fp32' % (B, W, H, S, S),
           'ms_per_batch': res[False], 'images_per_sec': B / (res[False] / 1e3), 'ms_per_batch_antialias': res[True],
           'launches_per_batch': 2, 'bound': 'hbm', 'algorithmic_bytes': int(alg), 'achieved': alg / (res[False] / 1e3) / 1e9,
           'peak': peak, 'unit': 'GB/s', 'frac': alg / (res[False] / 1e3) / 1e9 / peak}
    if with_cpu:
        torch.set_num_threads(1)
        t0 = time.time()
        ncpu = 4
        data_cpu.make_batch(frames[:ncpu], Ks[:ncpu], rs[:ncpu].cpu().numpy(), S)
        dt = (time.time() - t0) / ncpu
        torch.set_num_threads(os.cpu_count() or 1)
        fig['cpu_baseline'] = {'value': 1.0 / dt, 'unit': 'images/sec', 'cores': 1, 'kind': 'port',
                               'sample': '%d frames through oracle/data_cpu.py (the statements of Wild6DDataset.__getitem__ after the '
                                         'file reads: numpy bounding box + 3 torchvision resized_crop), one DataLoader worker' % ncpu}
    return fig


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    # exactly ONE line on stdout: libraries (NCCL's version banner, cuDNN notices) write to fd 1 as well -- park the real
    # stdout and point fd 1 at stderr until the JSON line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(line, flush=True)
    os.environ['SCP_VIT_PRECISION'] = args.vit_precision
    import torch
    import torch.distributed as dist
    from self_corr_pose_b200 import _lib
    _lib.lib()   # fail loudly if the native library is missing
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # keep stdout to the one JSON line (NCCL_DEBUG=VERSION)
        dist.init_process_group('nccl', init_method='env://', device_id=dev)
    if args.workload == 'config1' and args.batch == 64:
        args.batch = 32
    B = args.batch
    assert B % 4 == 0
    timed = Timer(torch, dist, dev, world)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    hp = None
    if args.workload == 'trainer':
        main_arm = trainer_arm(args, torch, dist, dev, world, rank, local, timed)
        if world == 1 and not args.no_kernel_breakdown:      # second reported figure + objects for the kernel breakdown
            hp = hotpath_arm(args, torch, dist, dev, world, rank, timed)
    elif args.workload == 'hotpath':
        main_arm = hp = hotpath_arm(args, torch, dist, dev, world, rank, timed)
    else:
        main_arm = config1_arm(args, torch, dev, timed)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    roof, kernels = (None, [])
    if hp is not None and not args.no_kernel_breakdown:
        roof, kernels = kernel_breakdown(torch, hp['hot'], hp['data'], hp['enc'], B, peaks)
    cpu = None
    if not args.no_cpu_baseline and world == 1 and args.workload != 'config1':
        cpu, _ = cpu_step_rate(args, dict(img_size=256, corr_h=64, corr_w=64), 'cpu_baseline')
    ms = main_arm['ms']
    line = {
        'metric': 'images/sec fwd+bwd (feat+corr+render+loss) 256x256', 'value': B * world / (ms / 1e3), 'unit': 'images/sec',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 (DINO ViT tensor-core products: %s)' % ('bf16x3 split = fp32-class, the parity mode'
                                                              if args.vit_precision == 'x3' else 'plain bf16, labelled FAST MODE outside the 1e-3 parity contract'),
        'data': 'synthetic (seeded synthetic DINO / ResNet weights: no checkpoints offline)',
        'config': {'workload': WORKLOAD_TEXT[args.workload] % (main_arm['N'], main_arm['nf']),
                   'images_per_gpu': B, 'img_size': 256, 'mesh': args.mesh, 'parallelism': 'dp%d' % world,
                   'vit_precision': args.vit_precision,
                   'l2': 'working set per step (>1 GB) exceeds the 126 MB L2; no explicit flush',
                   'cuda_graph': main_arm.get('graph', False)},
        'e2e': None if main_arm['ms_e2e'] is None else {
            'value': B * world / (main_arm['ms_e2e'] / 1e3), 'unit': 'images/sec', 'h2d_bytes_per_step': main_arm['h2d'],
            'd2h_bytes_per_step': main_arm['d2h'], 'ms_per_step': main_arm['ms_e2e']},
        'gpu_launches': main_arm['launches'] * args.steps,
        'clocks': sampler.summary() if sampler else None,
        'roofline': roof, 'kernels': kernels, 'cpu_baseline': cpu, 'loss': main_arm['loss'],
    }
    if main_arm.get('comm'):
        line['comm'] = main_arm['comm']
    if args.workload == 'trainer' and hp is not None:
        line['hotpath'] = {'value': B / (hp['ms'] / 1e3), 'unit': 'images/sec', 'ms_per_step': hp['ms'],
                           'e2e_value': B / (hp['ms_e2e'] / 1e3), 'cuda_graph': hp['graph'],
                           'workload': WORKLOAD_TEXT['hotpath'] % (hp['N'], hp['nf'])}
    if world == 1 and args.workload == 'trainer' and not args.no_kernel_breakdown:
        line['data_path'] = data_path_figure(torch, B, peaks, not args.no_cpu_baseline)
    emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
