import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Fixture file -> dict of torch tensors / python scalars."""
    data = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for key in data.files:
        arr = data[key]
        out[key] = arr.item() if arr.ndim == 0 else torch.from_numpy(arr)
    return out


SOLVER_CASES = [
    "ista_readme_fista", "ista_readme_plain", "ista_planted_200", "ista_randn_200",
    "ista_ragged", "ista_wide_d", "ista_earlystop", "ista_earlystop_plain",
    "ista_one_iter", "ista_big_alpha", "ista_warmstart",
    "encode_init_zero", "encode_init_ridge", "encode_init_transpose",
    # round 2: shapes that reach the k-blocked tcgen05 kernel, the notebook's dictionary, the two
    # remaining initialisers (z0 of the fixture is the reference's start code)
    "r2_blocked_128x512", "r2_blocked_64x512_plain", "r2_notebook_289x300",
    "r2_encode_init_lstsq", "r2_encode_init_unif",
]


CONV_CASES = ["conv_8x8_fista", "conv_8x8_plain", "conv_8x8_warmstart", "conv_3x3_auto_earlystop",
              "conv_8x8_pad3", "conv_4x4_stride2", "conv_4x4_stride2_pad1", "conv_3x3_pad1_auto"]


@pytest.fixture(scope="session")
def build_extension():
    """Make sure the in-tree .so exists (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as entry
    entry._load_build_ext().build()
    import lasso_b200
    return lasso_b200
