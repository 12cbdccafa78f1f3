"""PointNet++ backbones: mirror of the reference's network/models/backbones.py.

Same class names, constructor arguments, attribute / ``state_dict`` names (sa1..sa3, fp3..fp1,
conv1, bn1) and tensor shapes, built on ``hotrack_b200.pointnet_utils``.  ``PointNet2Msg_fast`` is
the backbone HandTrackNet uses (reference hand_network.py:54, backbones.py:74-133).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet_utils as pu
from .pointnet_utils import (PointNetFeaturePropagation, PointNetFeaturePropagation_fast, PointNetSetAbstraction,
                             PointNetSetAbstraction_fast, PointNetSetAbstractionMsg, PointNetSetAbstractionMsg_fast)

# configs/pointnet_config/pointnet2_camera_shallow1.yml of the reference, as a dict
SHALLOW1_CFG = {
    "sa1": {"npoint": 256, "radius_list": [0.1], "nsample_list": [32], "mlp_list": [[32, 32, 64]]},
    "sa2": {"npoint": 128, "radius_list": [0.2], "nsample_list": [32], "mlp_list": [[64, 64, 128]]},
    "sa3": {"mlp": [128, 128, 512]},
    "fp3": {"mlp": [256, 256]},
    "fp2": {"mlp": [256, 128]},
    "fp1": {"mlp": [128, 128]},
}


def default_cfg(device="cuda"):
    """The cfg dict HandTrackNet hands to its backbone (configs/config.py:48-54,93)."""
    return {"pointnet": {"camera": SHALLOW1_CFG}, "device": device,
            "network": {"handframe": "camera", "backbone_out_dim": 384}}


def _build_encoder(self, net_cfg, sa_msg, sa_all):
    self.sa1 = sa_msg(npoint=net_cfg["sa1"]["npoint"], radius_list=net_cfg["sa1"]["radius_list"],
                      nsample_list=net_cfg["sa1"]["nsample_list"], in_channel=self.in_dim + 3,
                      mlp_list=net_cfg["sa1"]["mlp_list"])
    self.sa2 = sa_msg(npoint=net_cfg["sa2"]["npoint"], radius_list=net_cfg["sa2"]["radius_list"],
                      nsample_list=net_cfg["sa2"]["nsample_list"], in_channel=self.sa1.out_channel + 3,
                      mlp_list=net_cfg["sa2"]["mlp_list"])
    self.sa3 = sa_all(npoint=None, radius=None, nsample=None, in_channel=self.sa2.out_channel + 3,
                      mlp=net_cfg["sa3"]["mlp"], group_all=True)


def _build_decoder(self, net_cfg, fp):
    self.fp3 = fp(in_channel=self.sa2.out_channel + self.sa3.out_channel, mlp=net_cfg["fp3"]["mlp"])
    self.fp2 = fp(in_channel=self.sa1.out_channel + self.fp3.out_channel, mlp=net_cfg["fp2"]["mlp"])
    self.fp1 = fp(in_channel=self.in_dim + 3 + self.fp2.out_channel, mlp=net_cfg["fp1"]["mlp"])
    self.conv1 = nn.Conv1d(self.fp1.out_channel, self.out_dim, 1)
    self.bn1 = nn.BatchNorm1d(self.out_dim)


class PointNet2Msg(nn.Module):
    """Reference backbones.py:17-71.  input (B,3+F,N) -> (B,out_dim,N)."""

    def __init__(self, cfg, out_dim, net_type="camera", use_xyz_feat=False, init_feature_dim=0):
        super().__init__()
        self.out_dim = out_dim
        self.in_dim = init_feature_dim + 3 if use_xyz_feat else init_feature_dim
        self.use_xyz_feat = use_xyz_feat
        _build_encoder(self, cfg["pointnet"][net_type], PointNetSetAbstractionMsg, PointNetSetAbstraction)
        _build_decoder(self, cfg["pointnet"][net_type], PointNetFeaturePropagation)
        self.device = cfg["device"]

    def forward(self, input):
        with pu.coord_scope():
            return self._forward(input)

    def _forward(self, input):
        l0_xyz = input[:, :3]
        l0_points = input if self.use_xyz_feat else input[:, 3:]
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        l0_points = self.fp1(l0_xyz, l1_xyz, torch.cat([l0_xyz, l0_points], dim=1), l1_points)
        return _head(self, l0_points)


class PointNet2Msg_fast(nn.Module):
    """Reference backbones.py:74-133: the part-batched variant HandTrackNet uses (P = 1).
    input (B,3+F,N) -> (B,out_dim,N)."""

    def __init__(self, cfg, out_dim, net_type="camera", use_xyz_feat=False, init_feature_dim=0):
        super().__init__()
        self.out_dim = out_dim
        self.in_dim = init_feature_dim + 3 if use_xyz_feat else init_feature_dim
        self.use_xyz_feat = use_xyz_feat
        _build_encoder(self, cfg["pointnet"][net_type], PointNetSetAbstractionMsg_fast, PointNetSetAbstraction_fast)
        _build_decoder(self, cfg["pointnet"][net_type], PointNetFeaturePropagation_fast)
        self.device = cfg["device"]
        # Fused engine: the encoder and FP3's first layer keep two-plane (hi + lo fp16) rows.  FP3 batch-normalises a
        # global feature broadcast over the points of a cloud -- nearly the same vector for every cloud -- so every
        # rounding made before that normalisation comes out ~40x larger at the end of the network (measured:
        # tools/dev/emul_prec.py); everything behind it is well conditioned and stays fp16.
        for m in (self.sa1, self.sa2, self.sa3):
            m.precise_layers = 99
        self.fp3.precise_layers = 1

    def forward(self, input):
        with pu.coord_scope():  # coordinate transposes shared by SA1 / SA2 / FP2 / FP1 (pointnet_utils.t_contig)
            return self._forward(input)

    def _forward(self, input):
        B, C, N = input.shape
        input = input.reshape(B, 1, C, N)
        l0_xyz = input[:, :, :3]
        l0_points = input if self.use_xyz_feat else input[:, :, 3:]
        # (only where SA1's MLP is long enough to hide them: at B=1 the fork / join costs more than it hides -- measured on
        # the tracked frame, B=1 x N=8192: 0.96 ms without, 1.17 ms with; training step at B=32 x N=4096: 3.05 -> 3.02 ms)
        prefetched = input.is_cuda and _PREFETCH and B * N >= 65536
        if prefetched:
            self._prefetch_searches(l0_xyz)
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points)
        if prefetched:  # SA1's MLP is longer than the prefetched searches: this join costs nothing and makes every later use safe
            torch.cuda.current_stream(input.device).wait_stream(_SEARCH_SIDE[input.device.index])
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        skip = torch.cat([l0_xyz, l0_points], dim=-2) if l0_points.shape[-2] else l0_xyz  # backbones.py:127-130
        # FP1's output goes to the fused head and nowhere else: its fp32 (B,C,N) copy need not be written
        self.fp1._rows_only = getattr(self.fp1, "engine", "ops") == "fused"
        l0_points = self.fp1(l0_xyz, l1_xyz, skip, l1_points)
        return _head(self, pu._carry(l0_points, l0_points.reshape(B, -1, N)))


_PREFETCH = __import__("os").environ.get("PN2_SEARCH_PREFETCH", "1") != "0"
_SEARCH_SIDE = {}  # device index -> side stream of the prefetched neighbour searches


def _prefetch_searches(self, l0_xyz):
    """Everything the backbone's later stages need from the COORDINATES alone -- SA2's sampling and ball query, FP2's and
    FP1's three-NN -- started on a side stream as soon as SA1's sampling has produced the level-1 coordinates, so that it
    runs beside SA1's MLP instead of between the stages (~70 us of the step's critical path at B=32, N=4096; FPS of 256
    points keeps 32 SMs busy for 17 us with nothing beside it).  Results land in the forward pass's memo
    (pointnet_utils.memo_call); sa2 / fp2 / fp1 find them there and wait for the event behind each."""
    B, P, C, _ = l0_xyz.shape
    cur = torch.cuda.current_stream(l0_xyz.device)
    side = _SEARCH_SIDE.get(l0_xyz.device.index)
    if side is None:
        side = _SEARCH_SIDE[l0_xyz.device.index] = torch.cuda.Stream(device=l0_xyz.device)
    xyz0, l1, _ = self.sa1.search(l0_xyz)  # on THIS stream: SA1 needs it at once (memo hit when sa1 runs)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        l1_4d = l1.unsqueeze(1).expand(B, P, C, l1.shape[-1])  # what sa1 returns as its first output
        xyz1, l2, _ = self.sa2.search(l1_4d)
        if getattr(self.fp1, "engine", "ops") == "fused":
            from . import fused
            fused.three_nn_sq(pu.t_contig(xyz0), pu.t_contig(xyz1))   # FP1: level 0 <- level 1
            fused.three_nn_sq(pu.t_contig(xyz1), pu.t_contig(l2))     # FP2: level 1 <- level 2


PointNet2Msg_fast._prefetch_searches = _prefetch_searches


def _head(self, feats):
    """relu(bn1(conv1(x))) (backbones.py:69,131-132); fused engine: one 16-bit tensor-core stack."""
    if getattr(self.fp1, "engine", "ops") == "fused":  # the engine the backbone was BUILT with (as every SA / FP module)
        from . import fused
        # rows_only_output (set by a caller that hands the features to fused gather stacks only -- HandTrackNet's q1 / q2):
        # the 201 MB fp32 channel-major copy of the output is then never written
        return fused.dense_stack(feats, [self.conv1], [self.bn1], self.training,
                                 rows_only=getattr(self, "rows_only_output", False))
    return F.relu(self.bn1(self.conv1(feats)))


class PointNet2Encoder(nn.Module):
    """Reference backbones.py:135-186: SA1-SA3 then a two-layer head on the global feature."""

    def __init__(self, cfg, out_dim, net_type="camera", use_xyz_feat=False, use_init_label=False, use_one_hot=False):
        super().__init__()
        self.out_dim = out_dim
        self.in_dim = (3 if use_xyz_feat else 0) + (1 if use_init_label else 0) + (22 if use_one_hot else 0)
        self.use_xyz_feat = use_xyz_feat
        _build_encoder(self, cfg["pointnet"][net_type], PointNetSetAbstractionMsg, PointNetSetAbstraction)
        self.conv1 = nn.Conv1d(self.sa3.out_channel, 256, 1)
        self.bn1 = nn.BatchNorm1d(256)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(256, self.out_dim, 1)
        self.bn2 = nn.BatchNorm1d(self.out_dim)
        self.device = cfg["device"]

    def forward(self, input):
        l0_xyz = input[:, :3]
        l0_points = input if self.use_xyz_feat else input[:, 3:]
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        _, l3_points = self.sa3(l2_xyz, l2_points)
        feat = self.drop1(F.relu(self.bn1(self.conv1(l3_points))))
        return F.relu(self.bn2(self.conv2(feat)))
