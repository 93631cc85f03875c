"""Per-kernel registers / stack (spill) / static shared memory of libscp_b200.so from `cuobjdump -res-usage`, as a
markdown table.   python tools/resource_usage.py > profiles/<round>_resource_usage.md   (no GPU needed)"""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'self_corr_pose_b200', 'libscp_b200.so')
txt = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True, check=True).stdout
rows, name = [], None
for line in txt.splitlines():
    m = re.match(r'\s*Function (\S+):', line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r'REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', line)
    if m and name:
        rows.append((name,) + tuple(int(x) for x in m.groups()))
        name = None
names = subprocess.run(['c++filt'], input='\n'.join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print('# Resource usage of every kernel in `libscp_b200.so` (sm_100a, `cuobjdump -res-usage`)\n')
print('STACK > 0 = bytes of per-thread local stack (register spills or indexed local arrays); dynamic shared memory is not '
      'listed here (GEMM: 4 x 48 KiB stages + staging, attention: see `scp_fa2.cuh`).\n')
print('| kernel | registers | stack B | static smem B |\n|---|---|---|---|')
for (_, reg, stack, shared, _local), n in sorted(zip(rows, names), key=lambda t: t[1]):
    n = re.sub(r'\(.*', '', n).replace('void ', '').replace('scp::', '')
    print('| `%s` | %d | %d | %d |' % (n, reg, stack, shared))
