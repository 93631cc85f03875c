"""Builds "R-GPU": the reference's own legacy SoftRas CUDA operator, recompiled for sm_100a.

Test / measurement infrastructure only -- the product never loads it.  The two reference source files
(third-party/softras/soft_renderer/cuda/soft_rasterize_cuda{.cpp,_kernel.cu}) are copied to a scratch directory under
/tmp, mechanically patched for the current torch API (three regexes: `.type().is_cuda()` -> `.is_cuda()`, `X.type()` ->
`X.scalar_type()`, `.data<T>()` -> `.data_ptr<T>()`; no arithmetic is touched) and compiled with
torch.utils.cpp_extension for compute_100a.  Only the built module `baseline/_ref/soft_rasterize.so` (the pybind module
name `soft_rasterize` is fixed by the reference's PYBIND11_MODULE statement) lands in the
repository tree (git-ignored, travels to the GPU box); no reference source is written into the repo.

    python baseline/build_ref_gpu.py            # needs /root/reference; ~90 s
"""
import os
import re
import shutil
import sys
import tempfile

REF = os.environ.get('SCP_REFERENCE_ROOT', '/root/reference')
SRC = os.path.join(REF, 'third-party', 'softras', 'soft_renderer', 'cuda')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
NAME = 'soft_rasterize_ref'


def patch(text):
    text = re.sub(r'\.type\(\)\.is_cuda\(\)', '.is_cuda()', text)
    text = re.sub(r'(\w+)\.type\(\)', r'\1.scalar_type()', text)
    text = re.sub(r'\.data<\s*([\w:]+)\s*>\(\)', r'.data_ptr<\1>()', text)
    return text


def main():
    if not os.path.isdir(SRC):
        sys.exit('reference tree not found at %s' % SRC)
    os.makedirs(OUT, exist_ok=True)
    os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    os.environ.setdefault('MAX_JOBS', '4')
    from torch.utils import cpp_extension
    tmp = tempfile.mkdtemp(prefix='scp_ref_gpu_')
    srcs = []
    for f in ('soft_rasterize_cuda.cpp', 'soft_rasterize_cuda_kernel.cu'):
        dst = os.path.join(tmp, f)
        with open(os.path.join(SRC, f)) as fi, open(dst, 'w') as fo:
            fo.write(patch(fi.read()))
        srcs.append(dst)
    bdir = os.path.join(tmp, 'build')
    os.makedirs(bdir)
    cpp_extension.load(name=NAME, sources=srcs, build_directory=bdir, verbose=True, is_python_module=False,
                       extra_cuda_cflags=['-O3'])
    so = os.path.join(bdir, NAME + '.so')
    shutil.copy(so, os.path.join(OUT, 'soft_rasterize.so'))
    shutil.rmtree(tmp, ignore_errors=True)
    print('built', os.path.join(OUT, 'soft_rasterize.so'))


if __name__ == '__main__':
    main()
