import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLDEN, name))
    d = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("sd/")}
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    return d, sd


TINY = dict(img_size=8, patch_size=2, in_chans=4, embed_dim=64, depth=2, num_heads=1, mlp_ratio=4,
            qkv_bias=False, mlp_time_embed=False, clip_dim=32, num_clip_token=5,
            enable_panoptic=True, use_ground_truth=False, num_panoptic_class=8)
