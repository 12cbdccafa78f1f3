"""HandTrackNet: mirror of the reference's network/models/hand_network.py:45-221 (the network around the pointnet_lib
path).  Same constructor (``cfg`` dict), attribute names and state_dict keys, same
``forward(input, flag_dict) -> ret_dict`` and ``compute_loss(input, ret_dict, flag_dict) -> (loss_dict, ret_dict)``
contracts, so the reference's Trainer / HandTrackModel drive it unchanged and its checkpoints load strictly.

What is done differently, outputs unchanged:
  * the backbone, q1 and q2 run on this package's engines (pointnet_utils.set_engine);
  * the hand frame (handframe == 'kp') and the rotation / translation losses use the GPU Kabsch kernel
    (hand_utils.solve_rot_and_trans -> csrc/kabsch.cu) instead of three CPU SVD round trips per step
    (reference hand_network.py:100,182-183 -> hand_utils.py:57-61);
  * the attention blocks are called with attn=False exactly as in the reference (:140-141) but do not evaluate the
    discarded multi-head attention (head_blocks.py), and TransT's point-cloud branch -- consumed only by that dead
    attention -- is skipped; the unused positional embedding (:122-125, "no use") is not computed either.
The MANO layer, IKNet and the visualisation hooks of the reference file are outside the path.
"""
import numpy as np
import torch
import torch.nn as nn

from .backbones import PointNet2Msg_fast
from .hand_utils import canonicalize, decanonicalize, handkp2palmkp, ransac_rt
from .head_blocks import PositionEmbeddingSine, TransT, attn_module, rearrange_module
from .pointnet_utils import PointNetSetAbstractionMsg_GivenCenterPoints, coord_scope, knn_point


def L2_loss(x, y, mask=None):  # x, y: [B,3,n], mask: [B,1,n]
    assert x.shape[1] == 3 and y.shape[1] == 3 and (mask is None or mask.shape[1] == 1)
    if mask is None:
        return (x - y).norm(dim=1).mean()
    return (((x - y) * mask).norm(dim=1).sum(dim=-1) / torch.clamp(mask.sum(dim=-1), min=1).squeeze()).mean()


def L1_loss(x, y, mask=None, check_dim_in=3):
    assert x.shape[1] == check_dim_in and y.shape[1] == check_dim_in and (mask is None or mask.shape[1] == 1)
    if mask is None:
        return (x - y).abs().mean()
    return (((x - y) * mask).abs().mean(dim=1).sum(dim=-1) / torch.clamp(mask.sum(dim=-1), min=1).squeeze()).mean()


def _rotation_angle_deg(rot):
    """mean geodesic angle of a batch of rotation matrices, in degrees (hand_network.py:210-218)."""
    trace = rot[:, 0, 0] + rot[:, 1, 1] + rot[:, 2, 2]
    return torch.mean(torch.acos(torch.clamp((trace - 1) / 2, min=-1, max=1))) * 180 / np.pi


class HandTrackNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.device = cfg['device']
        self.handframe = cfg['network']['handframe']
        c = cfg['network']['backbone_out_dim']
        assert c % 6 == 0
        self.bhand = PointNet2Msg_fast(cfg, c)
        # fused engine: the point features go to q1 / q2 (which read their row form) and to nothing else here -- TransT's
        # point-cloud branch is skipped --, so their fp32 (B,384,N) copy need not be written
        self.bhand.rows_only_output = getattr(self.bhand.fp1, "engine", "ops") == "fused"
        self.r1 = rearrange_module(channel=c)
        self.r2 = rearrange_module(channel=c)
        self.positionEmbedding = PositionEmbeddingSine(num_pos_feats=c // 6)
        mlps = [[128, 128, c // 2], [128, 128, c // 2]]
        self.q1 = PointNetSetAbstractionMsg_GivenCenterPoints(radius_list=[0.2, 0.2], nsample_list=[16, 64], mlp_list=mlps,
                                                              in_channel=c + 3, knn=True)
        self.q2 = PointNetSetAbstractionMsg_GivenCenterPoints(radius_list=[0.2, 0.2], nsample_list=[16, 64], mlp_list=mlps,
                                                              in_channel=c * 2 + 3, knn=True)
        self.transt = TransT(d_model=c)
        self.c3 = attn_module(d_model=c)
        self.final_mlp = nn.Sequential(nn.Conv1d(c, 256, 1), nn.ReLU(inplace=True), nn.Conv1d(256, 3, 1))

    # ---- hand frame -----------------------------------------------------------------------------------------------
    def canon_pose(self, input, jittered_kp, hand_points, palm_template):
        dev = hand_points.device
        if self.handframe == 'kp':
            rot, trans, _, _, _ = ransac_rt(palm_template, handkp2palmkp(jittered_kp))
            return {'scale': 0.2 * torch.ones(1, device=dev), 'rotation': rot, 'translation': trans}
        if self.handframe == 'OBB':
            return {k: v.to(dev).float() for k, v in input['OBB_pose'].items()}
        if self.handframe == 'camera':
            b = hand_points.shape[0]
            return {'scale': 0.2 * torch.ones(1, device=dev),
                    'rotation': torch.eye(3, device=dev).unsqueeze(0).repeat(b, 1, 1),
                    'translation': torch.zeros((b, 3, 1), device=dev)}
        raise NotImplementedError

    def forward(self, input, flag_dict):
        """input: jittered_hand_kp [B,21,3], hand_points [B,N,3] (+ the palm template) -> ret_dict with pred_kp [B,21,3]."""
        dev = self.device
        if flag_dict['track_flag']:
            palm_template = input['pred_palm_template']
        else:
            palm_template = input['gt_hand_pose']['palm_template'].to(dev)
        jittered_kp = input['jittered_hand_kp'].to(dev).float()
        hand_points = input['hand_points'].to(dev).float()
        canon = self.canon_pose(input, jittered_kp, hand_points, palm_template)
        ret = {'canon_pose': canon}
        kp_num = jittered_kp.shape[1]
        # one canonicalisation of the concatenation, as the reference (:118-119): the coordinates must be the same bits
        # for the index tensors to be the same
        cam = canonicalize(torch.cat([hand_points, jittered_kp], dim=1).transpose(2, 1), canon)
        xyz2, xyz1 = cam[..., :-kp_num], cam[..., -kp_num:]

        with coord_scope():  # the backbone and q1 share the point-major twin of xyz2
            src2 = self.bhand(xyz2)
            f11, group_idx = self.q1(xyz2, src2, xyz1, None, return_group_idx=True)
            f13 = self.q2(xyz2, src2, xyz1, self.r1(f11), pre_group_idx=group_idx)
        f14 = self.r2(f13)
        f15, _ = self.transt(src1=f14, pos1=None, src2=src2, pos2=None, attn=False, need_result2=False)
        fused = self.c3(f15, None, None, None, attn=False)
        ret['pred_kp_handframe'] = self.final_mlp(fused) + xyz1
        ret['init_kp_handframe'] = xyz1
        ret['points_handframe'] = xyz2
        ret['pred_kp'] = decanonicalize(ret['pred_kp_handframe'], canon).transpose(2, 1)
        assert ret['pred_kp'].shape[1] == kp_num
        if flag_dict.get('IKNet_flag'):
            ret['pred_kp_vis_mask'] = visibility_mask(ret['pred_kp'], hand_points)
        return ret

    def compute_loss(self, input, ret_dict, flag_dict):
        dev = self.device
        gt_kp = input['gt_hand_kp'].to(dev).float().transpose(-1, -2)     # [B,3,21]
        pred_kp = ret_dict['pred_kp'].transpose(-1, -2)
        canon = ret_dict['canon_pose']
        ret_dict['gt_kp_handframe'] = canonicalize(gt_kp, canon)
        scale = canon['scale'][:, None, None]
        init_s, pred_s, gt_s = (ret_dict[k] * scale for k in ('init_kp_handframe', 'pred_kp_handframe', 'gt_kp_handframe'))
        loss = {'hand_pred_kp_loss': L1_loss(pred_s, gt_s), 'hand_pred_kp_diff': L2_loss(pred_kp, gt_kp),
                'hand_init_kp_diff': L2_loss(init_s, gt_s)}
        if self.handframe != 'OBB':
            if 'global_pose' in ret_dict:
                gt_r = input['gt_hand_pose']['rotation'].to(dev).float().reshape(-1, 3, 3)
                gt_t = input['gt_hand_pose']['translation'].to(dev).float().reshape(-1, 3, 1)
                d_r = ret_dict['global_pose']['rotation'].reshape(-1, 3, 3)
                d_t = ret_dict['global_pose']['translation'].reshape(-1, 3, 1)
            else:
                palm = input['gt_hand_pose']['palm_template'].to(dev)
                gt_r, gt_t, _, _, _ = ransac_rt(palm, handkp2palmkp(gt_s.transpose(-1, -2)))
                d_r, d_t, _, _, _ = ransac_rt(palm, handkp2palmkp(pred_s.transpose(-1, -2)))
                loss['hand_init_r_diff'] = _rotation_angle_deg(gt_r)
                loss['hand_init_t_diff'] = gt_t.norm(dim=1).mean()
            loss['hand_pred_r_loss'] = L1_loss(d_r, gt_r)
            loss['hand_pred_t_loss'] = L1_loss(d_t, gt_t)
            loss['hand_pred_r_diff'] = _rotation_angle_deg(torch.matmul(d_r.transpose(-1, -2), gt_r))
            loss['hand_pred_t_diff'] = L2_loss(d_t, gt_t)
        if flag_dict['track_flag']:
            gt_rot = input['gt_hand_pose']['rotation'].to(dev).float().reshape(-1, 3, 3)
            gt_tr = input['gt_hand_pose']['translation'].to(dev).float().reshape(-1, 3, 1)
            loss['hand_canon_r_diff'] = _rotation_angle_deg(torch.matmul(canon['rotation'].reshape(-1, 3, 3).transpose(-1, -2), gt_rot))
            loss['hand_canon_t_diff'] = L2_loss(gt_tr, canon['translation'].reshape(-1, 3, 1))
        if flag_dict.get('IKNet_flag') and 'MANO_theta' in ret_dict:
            gt_theta = input['gt_hand_pose']['mano_pose'][:, 3:].to(dev).float()
            loss['MANO_theta_diff'] = L1_loss(ret_dict['MANO_theta'], gt_theta, check_dim_in=45)
        return loss, ret_dict


def visibility_mask(pred_kp, hand_points):
    """Joint visibility from the mean distance to the 4 nearest observed points (reference hand_network.py:149-155):
    [B,21,3], [B,N,3] -> bool [B,21]; the wrist and thumb-base thresholds are 1 cm looser."""
    d4, _ = knn_point(4, pred_kp, hand_points)
    d4 = torch.mean(d4, dim=-1)
    relax = torch.zeros(d4.shape[1], device=d4.device)
    relax[:2] = 0.01
    return (d4 - relax) < 0.02
