"""Process-level switches of the product that are NOT part of the reference's flag set."""
import os


def synthetic_weights_allowed():
    """No checkpoint / prior file can be downloaded offline.  The reference fails (torch.load / trimesh) or downloads
    (torchvision ImageNet weights) when a pretrained file is absent; so does this package -- UNLESS the caller opts in to
    seeded synthetic stand-ins with SCP_SYNTHETIC_WEIGHTS=1 (tests, bench.py and smoke() do; results obtained with
    stand-in weights say nothing about the trained model's accuracy, only about the kernels)."""
    return os.environ.get('SCP_SYNTHETIC_WEIGHTS', '0') not in ('', '0', 'false', 'False')


def pinned(t):
    """Page-locked copy of a host tensor when a CUDA driver is present (CPU-only test runs keep pageable memory)."""
    import torch
    return t.pin_memory() if torch.cuda.is_available() else t
