"""A/B of kernel variants on the GPU box: for the product library and every `libscp_b200.NAME.so` found next to it
(built here with `python -m self_corr_pose_b200.build --variant NAME -DMACRO=1`), run the parity tests that cover the
changed kernel and a timing script, each in its own process with SCP_LIB_VARIANT set.

    python tools/ab_variants.py --tests 'tests/test_vit_gpu.py' --time 'tools/time_vit.py 64' [--only early_qk]

Prints one line per variant: test verdict + the timing script's last output line.  A variant is only worth keeping when
its tests pass AND it is faster; the product default never changes here (the macro's default has to be flipped in the
header, rebuilt and re-verified)."""
import os as _os; _os.environ.setdefault("SCP_SYNTHETIC_WEIGHTS", "1")
import argparse
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument('--tests', default='tests/test_vit_gpu.py', help='pytest arguments (quoted), run with -m gpu -x -q')
ap.add_argument('--time', default='tools/time_vit.py 64', help='timing command (quoted), run with the same variant')
ap.add_argument('--only', default='', help='comma-separated variant names (default: all found)')
ap.add_argument('--timeout', type=int, default=600)
a = ap.parse_args()

found = sorted(re.match(r'libscp_b200\.(.+)\.so$', os.path.basename(p)).group(1)
               for p in glob.glob(os.path.join(ROOT, 'self_corr_pose_b200', 'libscp_b200.*.so')))
variants = [''] + [v for v in found if not a.only or v in a.only.split(',')]
for v in variants:
    env = dict(os.environ, SCP_LIB_VARIANT=v)
    if not v:
        env.pop('SCP_LIB_VARIANT')
    label = v or 'product'
    try:
        t = subprocess.run([sys.executable, '-m', 'pytest', '-m', 'gpu', '-x', '-q'] + a.tests.split(), cwd=ROOT, env=env,
                           capture_output=True, text=True, timeout=a.timeout)
        verdict = (t.stdout.strip().splitlines() or ['?'])[-1]
    except subprocess.TimeoutExpired:
        verdict = 'TIMEOUT after %d s' % a.timeout
    try:
        r = subprocess.run([sys.executable] + a.time.split(), cwd=ROOT, env=env, capture_output=True, text=True,
                           timeout=a.timeout)
        timing = (r.stdout.strip().splitlines() or [r.stderr.strip()[-200:]])[-1]
    except subprocess.TimeoutExpired:
        timing = 'TIMEOUT'
    print('%-12s | tests: %s | %s' % (label, verdict, timing), flush=True)
