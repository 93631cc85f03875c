"""Per-launch table of a `ncu --metrics ... --csv` log of tools/ncu_corr.py (gpurun_out/corr_ncu*.csv)."""
import collections
import csv
import sys

SHORT = {'gpu__time_duration.sum': 'ns', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
         'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor%',
         'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu%', 'smsp__inst_executed.sum': 'inst',
         'dram__bytes_read.sum': 'rdB', 'dram__bytes_write.sum': 'wrB', 'launch__registers_per_thread': 'regs',
         'launch__grid_size': 'grid',
         'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'longsb',
         'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'shortsb',
         'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'wait',
         'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'math',
         'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio': 'lg',
         'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio': 'mio',
         'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum': 'ldsect', 'lts__t_sectors_op_write.sum': 'l2wr',
         'lts__t_sectors_op_read.sum': 'l2rd', 'l1tex__data_pipe_lsu_wavefronts.sum': 'lsu_wf',
         'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active': 'lsu_wb%'}


def main(path, first=0, last=10):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    ix = {h: i for i, h in enumerate(rows[0])}
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((int(r[ix['ID']]), r[ix['Kernel Name']][:58]), {})[r[ix['Metric Name']]] = r[ix['Metric Value']]
    for (i, k), m in d.items():
        if not first <= i < last:
            continue

        def f(x):
            try:
                return '%.4g' % float(x.replace(',', ''))
            except ValueError:
                return x
        print(i, k, ' '.join('%s=%s' % (SHORT[n], f(v)) for n, v in m.items() if n in SHORT))


if __name__ == '__main__':
    main(sys.argv[1], *(int(a) for a in sys.argv[2:]))
