"""CPU: the SHIPPED SoftRas CUDA source (csrc/scp_softras.cu), compiled for the host by tools/emu (every CUDA thread an
OS thread, warp collectives and __syncthreads as barriers), against the C oracle -- default build and the compile-time
kernel candidates of DESIGN.md section 7 (two pixels per lane forward, face record in shared memory backward).
A functional check of traversal / indexing / reductions that needs no GPU; parity proper stays with the -m gpu tests."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('g++') is None, reason='no host C++ compiler')
def test_emulated_softras_kernels_and_candidates_match_the_oracle(capsys):
    sys.path.insert(0, os.path.join(ROOT, 'tools', 'emu'))
    import run_emu
    rc = run_emu.main()
    out = capsys.readouterr().out
    assert rc == 0 and 'EMU CHECK PASSED' in out, out
    assert all(name in out for name in ('default', 'fwd2px', 'facesmem', 'face16x2'))


@pytest.mark.skipif(shutil.which('g++') is None, reason='no host C++ compiler')
def test_emulated_loss_geometry_cycle_and_correspondence_kernels_match_reference_statements(capsys):
    """csrc/scp_loss.cu, scp_geom.cu, scp_cycle.cu, scp_corr.cu (its mma.sync / cp.async helpers in their host statement),
    scp_nhwc.cu and scp_data.cu compiled for the host and called through the C ABI with host pointers: values and gradients
    against the reference's op-by-op statements in fp64 (the references of the -m gpu tests), the torch operators the
    encoder-glue kernels replace, and -- for the data path -- the reference dataset class's golden batch (bit-identical)."""
    sys.path.insert(0, os.path.join(ROOT, 'tools', 'emu'))
    import run_emu_ops
    rc = run_emu_ops.main()
    out = capsys.readouterr().out
    assert rc == 0 and 'EMU OPS CHECK PASSED' in out, out
    assert 'nhwc glue' in out and 'data path' in out
