import os
import sys

import pytest

# no checkpoints offline: the tests run on seeded synthetic weights / the packaged prior meshes (explicit opt-in)
os.environ.setdefault('SCP_SYNTHETIC_WEIGHTS', '1')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _restore_global_numerics():
    """Trainer.define_model switches TF32 matmuls on process-wide (the reference environment's default); keep the tests
    independent of their order."""
    import torch
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved


@pytest.fixture
def deterministic_topk(monkeypatch):
    """Same tie-break among exactly equal cycle distances on the CPU oracle and the GPU product (torch.topk's own choice is
    implementation-defined and differs between devices)."""
    from oracle import corr as ocorr
    from self_corr_pose_b200.model.module import pretrained_corr as P
    monkeypatch.setattr(ocorr, 'select_topk', ocorr.stable_topk)
    monkeypatch.setattr(P, 'select_topk', ocorr.stable_topk)
