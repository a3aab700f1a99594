import json
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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """-> (blob dict of numpy arrays, model_cfg AttrDict, state dict of torch tensors)"""
    from mssvt_b200.config import AttrDict
    blob = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    with open(os.path.join(GOLDEN, name + ".cfg.json")) as f:
        cfg = AttrDict(json.load(f))
    state = {k[len("state/"):]: torch.from_numpy(v) for k, v in blob.items() if k.startswith("state/")}
    return blob, cfg, state


GOLDEN_CASES = ["s0_b2_n1200", "mixed_b3_n500"]
