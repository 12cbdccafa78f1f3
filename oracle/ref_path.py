"""The pointnet_lib hot path of HandTrackNet assembled from the REFERENCE's own classes (oracle/_ref/pyref, unmodified):
backbone -> q1 -> q2 exactly as network/models/hand_network.py builds and calls them (:54,61-69 constructors,
:130-134 forward; the rearrange_module between q1 and q2 is torch.nn outside pointnet_lib and left out on both arms).
TEST / BENCH INFRASTRUCTURE ONLY: bench.py's reference arm and cpu_baseline leg import this module and nothing of
hotrack_b200, so the reference process never loads libpn2b200.so.
"""
import math

import torch
import torch.nn as nn

from . import ref_modules


class RefPointPath(nn.Module):
    def __init__(self, device, cuda=True):
        super().__init__()
        rpu, rbb = ref_modules.load(cuda=cuda)
        cfg = ref_modules.handtracknet_cfg(device, "camera")
        c = cfg["network"]["backbone_out_dim"]
        self.bhand = rbb.PointNet2Msg_fast(cfg, c)
        mlps = [[128, 128, c // 2], [128, 128, c // 2]]
        self.q1 = rpu.PointNetSetAbstractionMsg_GivenCenterPoints(radius_list=[0.2, 0.2], nsample_list=[16, 64],
                                                                  mlp_list=mlps, in_channel=c + 3, knn=True)
        self.q2 = rpu.PointNetSetAbstractionMsg_GivenCenterPoints(radius_list=[0.2, 0.2], nsample_list=[16, 64],
                                                                  mlp_list=mlps, in_channel=c * 2 + 3, knn=True)

    def forward(self, xyz2, xyz1):
        src2 = self.bhand(xyz2)
        f11, idx = self.q1(xyz2, src2, xyz1, None, return_group_idx=True)
        f13 = self.q2(xyz2, src2, xyz1, f11, pre_group_idx=idx)
        return src2, f11, f13, idx


def xavier_init(module, seed=0):
    """xavier_normal(gain=sqrt(2)) conv weights, zero bias, as the reference trainer (trainer.py:20-40); the same
    stream of random numbers as hotrack_b200.handtrack_path.init_weights, so both arms start from equal weights."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d)):
            std = math.sqrt(2.0) * math.sqrt(2.0 / (m.weight.shape[1] + m.weight.shape[0]))
            with torch.no_grad():
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)
                if m.bias is not None:
                    m.bias.zero_()
