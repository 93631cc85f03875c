"""BASELINE.json configs[4] under ncu: ONE forward + ONE backward launch per cell between cudaProfilerStart/Stop.

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/sweep_ncu.csv \\
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \\
        -k regex:"forward_kernel|backward_face_kernel|corr_fwd_kernel|corr_bwd_rows_kernel" python tools/sweep_ncu.py run
    python tools/sweep_ncu.py table gpurun_out/sweep_ncu.csv profiles/r2_kernel_sweep_configs4_ncu.md

Cells (in launch order, B = 32): SoftRas soft-texture (sigma 1e-3) and depth (sigma 1e-4) renders at 64/128/256/512 px x
{642 V / 1280 F, 1280 V / 2556 F, 2562 V / 5120 F}; fused correspondence at P = 256 / 1024 / 4096 x N = 1280, C = 64."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = 32
SOFTRAS_CELLS = [(kind, size, mesh) for kind in ('softtex', 'depth') for size in (64, 128, 256, 512)
                 for mesh in ('ico642', 'uv1280', 'ico2562')]
CORR_CELLS = [16, 32, 64]
NF = {'ico642': 1280, 'uv1280': 2556, 'ico2562': 5120}


def run():
    import torch
    import torch.nn.functional as F
    from tests import _scenes
    from self_corr_pose_b200 import synthetic
    from self_corr_pose_b200.soft_renderer import functional as srf
    from self_corr_pose_b200.ops.corr_match import corr_match
    from self_corr_pose_b200.model.module.correspondence import make_meshgrid
    meshes = {'ico642': synthetic.icosphere(3), 'uv1280': synthetic.uv_sphere(), 'ico2562': synthetic.icosphere(4)}
    g = torch.Generator().manual_seed(0)
    rot, trans = synthetic.random_poses(B, g)
    jobs = []
    for kind, size, name in SOFTRAS_CELLS:
        v, f = meshes[name]
        fv, sv, ff = _scenes.screen_faces(v, f, rot, trans)
        tex = srf.face_vertices(_scenes.vertex_colors(sv), ff).cuda()
        fvd = fv.cuda().requires_grad_(True)
        kw = dict(image_size=size, texture_type='vertex', **_scenes.RENDER_CONFIGS[kind])
        jobs.append((fvd, tex, kw))
    N, C = 1280, 64
    cjobs = []
    for hf in CORR_CELLS:
        P = hf * hf
        a = F.normalize(torch.randn(B, C, P, device='cuda'), 2, 1).requires_grad_(True)
        m = F.normalize(torch.relu(torch.randn(B, N, C, device='cuda')), 2, -1).requires_grad_(True)
        md = (torch.rand(B, P, device='cuda') > 0.4).float()
        cjobs.append((a, m, md, torch.randn(B, N, 3, device='cuda'), make_meshgrid(hf, hf, 'cuda'), hf))
    for fvd, tex, kw in jobs[:1]:                     # one untimed call: library / module initialisation
        srf.soft_rasterize(fvd, tex, **kw).sum().backward()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for fvd, tex, kw in jobs:
        o = srf.soft_rasterize(fvd, tex, **kw)
        torch.autograd.grad(o, fvd, torch.randn_like(o))
    for a, m, md, pv, grid, hf in cjobs:
        _, pool, mt, im, _A = corr_match(a, m, md, pv, grid, 10.0, hf, hf, want_full=False, want_pool=True)
        torch.autograd.grad([pool, mt, im], [a, m], [torch.randn_like(pool), torch.randn_like(mt), torch.randn_like(im)])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def table(src, dst):
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    hbm = peaks['hbm_gbs']
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi, ui, idi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
    per = {}
    order = []
    for r in rows[1:]:
        key = r[idi]
        if key not in per:
            per[key] = {'name': r[ki]}
            order.append(key)
        val = float(r[vi].replace(',', ''))
        unit = r[ui]
        scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
        per[key][r[mi]] = val * scale
    launches = [per[k] for k in order]
    pick = lambda pat: [l for l in launches if pat in l['name']]
    fw, bw = pick('forward_kernel'), pick('backward_face_kernel')
    cf, cb = pick('corr_fwd_kernel'), pick('corr_bwd_rows_kernel')

    def cols(l, alg_bytes):
        ms = l['gpu__time_duration.sum']
        dram = l['dram__bytes_read.sum'] + l['dram__bytes_write.sum']
        return '%.3f | %.1f | %.1f | %.4f | %.0f | %.0f | %.1f' % (
            ms, alg_bytes / ms / 1e6, dram / ms / 1e6, alg_bytes / ms / 1e6 / hbm,
            l['smsp__issue_active.avg.pct_of_peak_sustained_active'],
            l['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'],
            l['smsp__thread_inst_executed_per_inst_executed.ratio'])
    head = 'ms | algorithmic GB/s | DRAM GB/s | frac of HBM peak | issue slots % | tensor pipe % | lanes / instr'
    out = ['# BASELINE configs[4] under ncu (round 2 kernels), B = %d, one launch per cell, `--clock-control none`; HBM peak %.0f GB/s (measured)\n' % (B, hbm),
           'Per launch: duration, algorithmic bytes (SURVEY 8d) / duration, DRAM bytes actually moved / duration, the byte-roofline fraction,',
           'and what actually bounds these kernels: issue-slot and tensor-pipe utilisation, active lanes per instruction.  ncu times are',
           'cold-cache single launches (the CUDA-event sweep of warm launches is `r2_kernel_sweep_configs4.md`).\n',
           '## SoftRas forward (`forward_kernel`, without the pack pre-pass)\n', '| render | px | F | ' + head + ' |', '|---|---:|---:|' + '---:|' * 7]
    assert len(fw) == len(SOFTRAS_CELLS) and len(bw) == len(SOFTRAS_CELLS), (len(fw), len(bw))
    for (kind, size, mesh), l in zip(SOFTRAS_CELLS, fw):
        out.append('| %s | %d | %d | %s |' % (kind, size, NF[mesh], cols(l, B * (72 * NF[mesh] + 24 * size * size))))
    out += ['', '## SoftRas backward (`backward_face_kernel`)\n', '| render | px | F | ' + head + ' |', '|---|---:|---:|' + '---:|' * 7]
    for (kind, size, mesh), l in zip(SOFTRAS_CELLS, bw):
        out.append('| %s | %d | %d | %s |' % (kind, size, NF[mesh], cols(l, B * (144 * NF[mesh] + 40 * size * size))))
    N, C = 1280, 64
    out += ['', '## Fused correspondence, N = 1280, C = 64 (training variant: pooled pointcorr)\n', '| kernel | P | ' + head + ' |', '|---|---:|' + '---:|' * 7]
    for hf, l in zip(CORR_CELLS, cf):
        P = hf * hf
        out.append('| corr_fwd_kernel | %d | %s |' % (P, cols(l, B * (4 * (C * P + N * C + P + 3 * N) + 4 * (P * N // 4 + 2 * N + 3 * P)))))
    for hf, l in zip(CORR_CELLS, cb):
        P = hf * hf
        out.append('| corr_bwd_rows_kernel | %d | %s |' % (P, cols(l, B * 4 * (2 * C * P + 2 * N * C + P * N // 4 + 6 * P + 7 * N))))
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))


if __name__ == '__main__':
    if sys.argv[1] == 'run':
        run()
    else:
        table(sys.argv[2], sys.argv[3])
