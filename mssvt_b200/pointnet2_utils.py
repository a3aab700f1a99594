"""The four operators of pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py that the MsSVT
backbone calls, on top of libmssvt_b200.so: farthest_point_sample, gather_operation, three_nn,
grouping_operation.  Same names, signatures and results (including the reference FPS tie order
and the three-NN FMA contraction).  CUDA tensors only.
"""
from typing import Tuple

import torch
from torch.autograd import Function

from ._lib import call, ptr, stream


def _i32(t):
    return t if t.dtype == torch.int32 and t.is_contiguous() else t.to(torch.int32).contiguous()


class FarthestPointSampling(Function):
    """pointnet2_utils.py:10-36."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        output = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        # the running min-distance lives on chip; the (B, N) scratch of the reference is only
        # needed for rows that do not fit shared memory
        temp = torch.empty((B, N), dtype=torch.float32, device=xyz.device) if N > 50000 else None
        call("mssvt_fps", B, N, npoint, ptr(xyz), ptr(temp), ptr(output), stream())
        return output

    @staticmethod
    def backward(xyz, a=None):
        return None, None


farthest_point_sample = furthest_point_sample = FarthestPointSampling.apply


class GatherOperation(Function):
    """pointnet2_utils.py:39-73: (B, C, N), (B, npoint) -> (B, C, npoint)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous() and idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        call("mssvt_gather_points", B, C, N, npoint, ptr(features), ptr(_i32(idx)), ptr(output), stream())
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        # gather_points_grad == group_points_grad with one sample per point
        grad = torch.empty((B, C, N), dtype=torch.float32, device=grad_out.device)
        call("mssvt_group_points_grad", B, C, N, npoint, 1, ptr(grad_out.contiguous()), ptr(_i32(idx)),
             ptr(grad), stream())
        return grad, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """pointnet2_utils.py:76-105: returns (sqrt(dist2), idx)."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert unknown.is_contiguous() and known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
        call("mssvt_three_nn", B, N, m, ptr(unknown), ptr(known), ptr(dist2), ptr(idx), stream())
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class GroupingOperation(Function):
    """pointnet2_utils.py:156-197: (B, C, N), (B, npoint, nsample) -> (B, C, npoint, nsample)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous() and idx.is_contiguous()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = torch.empty((B, C, nfeatures, nsample), dtype=torch.float32, device=features.device)
        call("mssvt_group_points", B, C, N, nfeatures, nsample, ptr(features), ptr(_i32(idx)),
             ptr(output), stream())
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad = torch.empty((B, C, N), dtype=torch.float32, device=grad_out.device)
        call("mssvt_group_points_grad", B, C, N, npoint, nsample, ptr(grad_out.contiguous()),
             ptr(_i32(idx)), ptr(grad), stream())
        return grad, None


grouping_operation = GroupingOperation.apply
