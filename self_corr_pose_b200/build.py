"""Builds the C-ABI shared library `libscp_b200.so` (hand-written sm_100a CUDA) in-tree.

    python -m self_corr_pose_b200.build [--force]
    python -m self_corr_pose_b200.build --variant NAME -DMACRO=1 [-DMACRO2=...]     # kernel experiments only

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box
with the gpurun snapshot.

A VARIANT is the same library compiled with extra -D macros (the compile-time switches documented in the kernel
headers, e.g. SCP_FA2_EARLY_QK) into `libscp_b200.NAME.so`, loaded instead of the product library when the
environment variable SCP_LIB_VARIANT=NAME is set (`_lib.py`) -- for A/B runs of a candidate kernel on the GPU box
(`tools/ab_variants.py`); nothing in the product, the tests or bench.py sets that variable.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libscp_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    d.append(os.path.join(HERE, '..', 'include', 'scp_b200.h'))
    return d


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def lib_path(variant=None):
    return LIB if not variant else os.path.join(HERE, 'libscp_b200.%s.so' % variant)


def _compile(src, force, variant=None, defines=()):
    obj = os.path.join(CSRC, 'build', variant or '', os.path.basename(src)[:-3] + '.o')
    os.makedirs(os.path.dirname(obj), exist_ok=True)
    if force or _stale(obj, [src] + _deps()):
        cmd = [NVCC] + ARCH + FLAGS + list(defines) + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + '.log', 'w') as f:
            f.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stdout + r.stderr))
    return obj


def build(force=False, verbose=False, variant=None, defines=()):
    srcs = sources()
    out = lib_path(variant)
    force = force or bool(variant)          # a variant's macros are not tracked by the staleness check
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, variant, defines), srcs))
    if force or _stale(out, objs):
        cmd = [NVCC] + ARCH + ['-shared', '-o', out] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    if verbose:
        for o in objs:
            print(open(o + '.log').read())
    return out


if __name__ == '__main__':
    variant = sys.argv[sys.argv.index('--variant') + 1] if '--variant' in sys.argv else None
    defines = [a for a in sys.argv[1:] if a.startswith('-D')]
    if defines and not variant:
        sys.exit('-D macros build a variant: pass --variant NAME (the product library takes no macros)')
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, variant=variant, defines=defines))
