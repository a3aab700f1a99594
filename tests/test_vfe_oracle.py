"""CPU: the DynamicVFE restatement (oracle/vfe.py) against the golden vectors written from the reference's
unmodified module (oracle/pin_vfe_against_reference.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import vfe as orc_vfe

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["vfe_s0_b2_p6000", "vfe_two_layer_b3_p4000"])
def test_vfe_oracle_reproduces_reference_golden(name):
    blob = np.load(os.path.join(GOLDEN, name + ".npz"))
    state = {k[6:]: torch.from_numpy(blob[k]) for k in blob.files if k.startswith("state/")}
    f, c = orc_vfe.dynamic_vfe_forward(state, torch.from_numpy(blob["points"]), int(blob["batch_size"]),
                                       blob["voxel_size"].tolist(), blob["grid_size"].tolist(),
                                       blob["pc_range"].tolist(), 5)
    assert torch.equal(c, torch.from_numpy(blob["voxel_coords"]))
    ref = torch.from_numpy(blob["voxel_features"])
    assert (f - ref).abs().max().item() <= 1e-6 * max(ref.abs().max().item(), 1.0)
    # rows are unique voxels in ascending (b, x, y, z) order
    key = ((c[:, 0].long() * 1000 + c[:, 3]) * 1000 + c[:, 2]) * 1000 + c[:, 1]
    assert bool((key[1:] > key[:-1]).all())
