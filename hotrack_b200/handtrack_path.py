"""The pointnet_lib hot path of HandTrackNet as one module: backbone -> q1 -> q2.

Mirrors the slice of the reference's ``HandTrackNet`` that runs through pointnet_lib
(network/models/hand_network.py:54,61-69 constructors; :130-134 forward), with the same attribute
names (``bhand``, ``q1``, ``q2``) so the corresponding ``state_dict`` entries of a HandTrackNet
checkpoint load with ``strict=False``.  What lies between q1 and q2 in the full network -- the
``rearrange_module`` 1920->384 Conv1d (blocks.py:226-239) -- is torch.nn outside pointnet_lib and is
replaced by the identity here; the dead attention blocks, the MANO layer and the SVD hand frame are
out of scope.

``impl`` selects whose classes build the module: this package's (default) or the reference's own
(an object with ``PointNet2Msg_fast`` and ``PointNetSetAbstractionMsg_GivenCenterPoints``), which is
how the parity tests and bench.py's reference arm run the SAME path on the reference.
"""
import types

import torch
import torch.nn as nn

_SIDE = {}  # device index -> side stream for the early neighbour search (kept off the module: modules get deep-copied)


def _ours():
    from . import backbones, pointnet_utils

    return types.SimpleNamespace(
        PointNet2Msg_fast=backbones.PointNet2Msg_fast,
        PointNetSetAbstractionMsg_GivenCenterPoints=pointnet_utils.PointNetSetAbstractionMsg_GivenCenterPoints)


class HandTrackPointPath(nn.Module):
    def __init__(self, cfg, impl=None):
        super().__init__()
        impl = impl or _ours()
        c = cfg["network"]["backbone_out_dim"]
        self.bhand = impl.PointNet2Msg_fast(cfg, c)
        mlps = [[128, 128, c // 2], [128, 128, c // 2]]
        self.q1 = impl.PointNetSetAbstractionMsg_GivenCenterPoints(
            radius_list=[0.2, 0.2], nsample_list=[16, 64], mlp_list=mlps, in_channel=c + 3, knn=True)
        self.q2 = impl.PointNetSetAbstractionMsg_GivenCenterPoints(
            radius_list=[0.2, 0.2], nsample_list=[16, 64], mlp_list=mlps, in_channel=c * 2 + 3, knn=True)

    def forward(self, xyz2, xyz1):
        """xyz2 (B,3,N) canonicalised hand cloud, xyz1 (B,3,21) canonicalised joints ->
        (src2 (B,384,N), f11 (B,384,21), f13 (B,384,21), group indices)."""
        if not hasattr(self.q1, "group_indices"):  # the reference's classes (parity tests, bench.py's reference arm)
            return self._forward(xyz2, xyz1)
        from . import pointnet_utils as pu

        with pu.coord_scope():  # one forward pass: the (B,N,3) twins of the coordinate tensors are made once
            # ... and on THIS stream, before the side stream forks off (it reads them too)
            pu.t_contig(xyz2)
            pu.t_contig(xyz1)
            return self._forward(xyz2, xyz1)

    def _forward(self, xyz2, xyz1):
        pre = None
        if xyz2.is_cuda and hasattr(self.q1, "group_indices"):
            # q1's neighbour search needs the coordinates only: it runs on a side stream while the backbone starts
            # (farthest point sampling keeps 32 of the 148 SMs busy for ~100 us and nothing else can run beside it)
            side = _SIDE.get(xyz2.device.index)
            if side is None:
                side = _SIDE[xyz2.device.index] = torch.cuda.Stream(device=xyz2.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                pre = self.q1.group_indices(xyz2, xyz1)
        src2 = self.bhand(xyz2)
        if pre is not None:
            torch.cuda.current_stream().wait_stream(side)
        f11, idx = self.q1(xyz2, src2, xyz1, None, pre_group_idx=pre, return_group_idx=True)
        f13 = self.q2(xyz2, src2, xyz1, f11, pre_group_idx=idx)
        return src2, f11, f13, idx


def init_weights(module, seed=0):
    """xavier_normal(gain=sqrt(2)) conv weights, zero bias, as the reference trainer (trainer.py:20-40)."""
    import math

    import torch

    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d)):
            fan_in = m.weight.shape[1]
            fan_out = m.weight.shape[0]
            std = math.sqrt(2.0) * math.sqrt(2.0 / (fan_in + fan_out))
            with torch.no_grad():
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)
                if m.bias is not None:
                    m.bias.zero_()
