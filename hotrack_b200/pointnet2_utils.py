"""Operator API: mirror of the reference's pointnet_lib/pointnet2_utils.py.

Same seven ``autograd.Function``s (``.apply`` aliases with the reference's names,
argument order, tensor shapes and dtypes) and the three grouper modules, backed
by the sm_100a kernels through ``hotrack_b200.pointnet2_cuda``.  Reference
lines are cited per class (paths relative to
/root/reference/network/models/pointnet_lib/).

Differences that do not change results: outputs are allocated on the inputs'
device with ``torch.empty`` (the reference uses ``torch.cuda.IntTensor(...)`` on
the current device), FPS passes no ``temp`` scratch (the kernel keeps the running
distances in registers), and backward functions take ``ctx`` first everywhere
(the reference's no-grad Functions have malformed ``backward`` signatures that
are never called).
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointnet2_cuda as pointnet2


class FurthestPointSampling(Function):
    """pointnet2_utils.py:11-38.  xyz (B,N,3) -> (B,npoint) int32, first index 0."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        xyz = xyz.contiguous()
        B, N, _ = xyz.size()
        output = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        # up to 8192 points the running distances stay in registers; larger clouds stream them through `temp`
        temp = None if N <= 8192 else torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        pointnet2.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, output)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    """pointnet2_utils.py:41-77.  features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        features = features.contiguous()
        idx = idx.contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty(B, C, npoint, dtype=torch.float32, device=features.device)
        pointnet2.gather_points_wrapper(B, C, N, npoint, features, idx, output)
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        pointnet2.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class KNN(Function):
    """pointnet2_utils.py:79-109.  k, unknown (B,N,3), known (B,M,3) -> (dist (B,N,k), idx int32)."""

    @staticmethod
    def forward(ctx, k: int, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        unknown = unknown.contiguous()
        known = known.contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, k, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, N, k, dtype=torch.int32, device=unknown.device)
        pointnet2.knn_wrapper(B, N, m, k, unknown, known, dist2, idx)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None


knn = KNN.apply


class ThreeNN(Function):
    """pointnet2_utils.py:111-142.  unknown (B,N,3), known (B,M,3) -> (dist (B,N,3), idx int32)."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        unknown = unknown.contiguous()
        known = known.contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty(B, N, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, N, 3, dtype=torch.int32, device=unknown.device)
        pointnet2.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """pointnet2_utils.py:145-193.  features (B,C,M), idx (B,n,3) int32, weight (B,n,3) -> (B,C,n);
    gradient flows to ``features`` only."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        features = features.contiguous()
        idx = idx.contiguous()
        weight = weight.contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
        pointnet2.three_interpolate_wrapper(B, c, m, n, features, idx, weight, output)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.zeros(B, c, m, dtype=torch.float32, device=grad_out.device)
        pointnet2.three_interpolate_grad_wrapper(B, c, n, m, grad_out.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """pointnet2_utils.py:196-239.  features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        features = features.contiguous()
        idx = idx.contiguous().int()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = torch.empty(B, C, nfeatures, nsample, dtype=torch.float32, device=features.device)
        pointnet2.group_points_wrapper(B, C, N, nfeatures, nsample, features, idx, output)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        pointnet2.group_points_grad_wrapper(B, C, N, npoint, nsample, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """pointnet2_utils.py:242-272.  radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3) -> (B,npoint,nsample) int32."""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        new_xyz = new_xyz.contiguous()
        xyz = xyz.contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        # the kernel writes every slot (index 0 for an empty ball: what the reference's zero-initialised buffer keeps)
        idx = (torch.empty if N > 0 else torch.zeros)(B, npoint, nsample, dtype=torch.int32, device=xyz.device)
        pointnet2.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:275-310: ball query, group xyz (centre-relative) and features."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,nsample)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        if self.use_xyz:
            return torch.cat([grouped_features, grouped_xyz], dim=1)  # (B,C+3,npoint,nsample)
        return grouped_features


class GroupAll(nn.Module):
    """pointnet2_utils.py:313-336: one group holding every point, [xyz, features] order."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)  # (B,3+C,1,N)
        return grouped_features


class KNNAndGroup(nn.Module):
    """pointnet2_utils.py:339-385: kNN grouping, [xyz, features] order.

    The reference calls ``knn(xyz, new_xyz, self.radius, self.nsample)`` (:362), which
    does not match KNN.forward(k, unknown, known) and raises if ever reached; this
    mirror makes the evident intent work: the ``nsample`` nearest ``xyz`` of each ``new_xyz``.
    """

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor = None, idx: torch.Tensor = None,
                features: torch.Tensor = None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            _, idx = knn(self.nsample, new_xyz, xyz)  # (B,M,K)
        idx = idx.detach()
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,M,K)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)  # (B,3+C,M,K)
        return grouped_features
