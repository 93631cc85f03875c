"""`ncu -i rep --page raw --csv` -> markdown summary (one table per kernel instance of interest) + traffic json.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/summarize_ncu.py /tmp/raw.csv profiles/rN_hot_kernels_ncu_full.md profiles/traffic.json "<note>"
Only kernels of this package are listed (first two instances of each name)."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import csv
import json
import re
import sys

src, dst_md, dst_json = sys.argv[1:4]
note = sys.argv[4] if len(sys.argv) > 4 else ''
KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
scale = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0}
tscale = {'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6, 'msecond': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'second': 1e6}
seen = {}
out, traffic = [], {}
for part in src.split(','):            # several captures of the same step (different -k filters)
    rows = list(csv.reader(open(part)))
    hdr, units = rows[0], rows[1]      # second row holds the units
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[col['Kernel Name']])
        name = name.replace('void ', '').replace('scp::', '')
        if not re.search(r'softras::|corr::|corr_tc::|gemm_rs::|gemm::|vit::|fa[23]?::|loss::|geom::|cycle::|sym::|jitter::|nhwc::|data::|posefit::', name):
            continue
        k = seen.get(name, 0)
        seen[name] = k + 1
        if k >= 2:
            continue
        key = '%s #%d' % (name, k)
        out.append('## `%s`\n\n| metric | value |\n|---|---|' % key)
        rd = wr = us = 0.0
        extra = {}
        for m in KEEP:
            if m not in col:
                continue
            val, unit = r[col[m]], units[col[m]]
            out.append('| %s | %s %s |' % (m, val, unit))
            try:
                x = float(val.replace(',', ''))
            except ValueError:
                continue
            if m == 'dram__bytes_read.sum':
                rd = x * scale.get(unit, 1.0)
            elif m == 'dram__bytes_write.sum':
                wr = x * scale.get(unit, 1.0)
            elif m == 'gpu__time_duration.sum':
                us = x * tscale.get(unit, 1.0)
            elif m == 'smsp__issue_active.avg.pct_of_peak_sustained_active':
                extra['issue_active_pct'] = x
            elif m == 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active':
                extra['tensor_pipe_pct'] = x
            elif m == 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed':
                extra['dram_throughput_pct'] = x
        out.append('')
        traffic[key] = dict({'dram_bytes': rd + wr, 'us': us}, **extra)
open(dst_md, 'w').write('# `ncu --set full --clock-control none` of the hot kernels (one step, B = 64, 1xB200)\n\n%s\n\n' % note +
                        '\n'.join(out) + '\n')
json.dump(traffic, open(dst_json, 'w'), indent=1)
print('\n'.join('%-60s %10.1f us %8.1f MB' % (k, v['us'], v['dram_bytes'] / 1e6) for k, v in traffic.items()))
