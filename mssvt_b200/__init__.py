"""mssvt_b200 -- B200-native (sm_100a) implementation of the MsSVT mixed-scale sparse voxel
attention backbone hot path, behind the reference's own module and operator API.

  mssvt_b200.mssvt_backbone   MixedScaleSparseTransformer{,Block,CompressBlock}  (pcdet/models/backbones_3d/mssvt_backbone.py)
  mssvt_b200.mssvt_utils      SparseTensor, MixedScaleAttention                  (pcdet/models/model_utils/mssvt_utils.py)
  mssvt_b200.mssvt_ops        build_hash_table, get_non_empty_window_center, ...  (pcdet/ops/mssvt/mssvt_ops.py)
  mssvt_b200.pointnet2_utils  farthest_point_sample, gather_operation, three_nn,  (pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py)
                              grouping_operation
  mssvt_b200.csrc             hand-written CUDA kernels + the C-ABI of include/mssvt_b200.h

Importing the package needs neither a GPU nor the built library; the first operator call loads
libmssvt_b200.so and raises if it is missing (there is no CPU or PyTorch fallback).
"""
from .config import AttrDict, block_cfg, compress_cfg, s0_model_cfg  # noqa: F401

__version__ = "0.1.0"
