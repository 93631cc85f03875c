"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-step table.

    python tools/summarize_launches.py gpurun_out/launches.csv.gz profiles/rN_bench_launch_summary.md "<note>"

The step count is recovered from the attention kernel (nine launches per step)."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import collections
import csv
import gzip
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ''
op = gzip.open if src.endswith('.gz') else open
rows = list(csv.reader(l for l in op(src, 'rt') if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki])[:110]
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(',', '')) * scale[r[ui]]
steps = max(v[0] for k, v in agg.items() if re.search(r'fa[23]?::fa[23]?_fwd_kernel', k)) // 9
total = sum(v[1] for v in agg.values()) / steps
mine = sum(v[1] for k, v in agg.items() if re.search(r'softras::|corr::|corr_tc::|gemm_rs::|gemm::|vit::|fa[23]?::|loss::|geom::|cycle::|sym::|jitter::|nhwc::|data::|posefit::', k)) / steps
with open(dst, 'w') as f:
    f.write('# ncu launch list of `bench.py`, aggregated per step\n\n%s\n\n' % note)
    f.write('%d launches over %d steps (warm-up, timed and end-to-end steps all run under the profiler); times are '
            'cold-cache and serialised: compare SHARES.\n\n' % (len(rows) - 1, steps))
    f.write('Sum of kernel time per step: %.2f ms.  This package\'s kernels: %.2f ms (%.1f %%); torch/ATen glue: '
            '%.2f ms (%.1f %%).\n\n' % (total, mine, 100 * mine / total, total - mine, 100 * (total - mine) / total))
    f.write('| ms/step | launches/step | share | kernel |\n|---:|---:|---:|---|\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
        f.write('| %.3f | %.1f | %.1f %% | `%s` |\n' % (v[1] / steps, v[0] / steps, 100 * v[1] / steps / total, k))
print(open(dst).read()[:1500])
