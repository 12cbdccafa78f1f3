"""The blocks of HandTrackNet's head around the pointnet_lib path: ``rearrange_module``
(reference network/models/blocks.py:226-239), ``attn_module`` / ``TransT`` / ``PositionEmbeddingSine``
(network/models/transformer.py:17-123 -- "now this file is used as a high-performance MLP", :1-4).  Constructor
arguments, attribute names and state_dict keys are the reference's (MultiheadAttention weights included), so
HandTrackNet checkpoints load strictly.

What differs: with ``attn=False`` -- how HandTrackNet calls every block (hand_network.py:140-141) -- the reference still
EVALUATES the multi-head attention and then discards it (transformer.py:78-82, ``src1_new = src1``); for the point-cloud
self-attention of ``TransT.s12`` that is 8 heads x B x N x N scores (17 GB at B=32, N=4096) and ~0.8 TFLOP of dead fp32
GEMMs.  Here attention is evaluated only when ``attn=True``; the outputs are the same tensors.  (In training mode the
skipped blocks draw no dropout masks, so the live feed-forward dropouts see a different random stream than the
reference's for the same seed; their distribution is unchanged.)
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

# kinematic neighbours of the 21 hand joints (reference blocks.py:229-232): next joint towards the tip, parent joint, and
# the corresponding joints of the two adjacent fingers
_NEIGHBOURS = (
    (1, 2, 3, 4, 4, 6, 7, 8, 8, 10, 11, 12, 12, 14, 15, 16, 16, 18, 19, 20, 20),
    (17, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 0, 13, 14, 15, 0, 17, 18, 19),
    (1, 1, 2, 3, 4, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16),
    (17, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 17, 18, 19, 20),
)


class rearrange_module(nn.Module):
    """[B, C, 21] -> [B, C, 21]: each joint's feature stacked with those of its four neighbours, 1x1 conv 5C -> C."""

    def __init__(self, channel=384, add_points=False, re=5):
        super().__init__()
        self.rearrange1, self.rearrange2, self.rearrange3, self.rearrange4 = (list(t) for t in _NEIGHBOURS)
        self.re = re
        self.linear = nn.Conv1d(channel * re, channel, 1)
        # gather indices on the device once (not in state_dict): an index list would be a host->device copy per call
        self.register_buffer("_gather", torch.tensor(_NEIGHBOURS, dtype=torch.long).reshape(-1), persistent=False)

    def forward(self, new_points):
        B, C, J = new_points.shape
        nb = new_points.index_select(2, self._gather).view(B, C, len(_NEIGHBOURS), J)       # [B,C,4,21]
        stacked = torch.cat([new_points.unsqueeze(2), nb], dim=2).transpose(1, 2)            # [B,5,C,21]
        return self.linear(stacked.reshape(B, (len(_NEIGHBOURS) + 1) * C, J))


_ACTIVATIONS = {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}


class attn_module(nn.Module):
    """(optional attention) -> LayerNorm -> (feed-forward d -> dim_feedforward -> d, residual, LayerNorm unless no_linear)."""

    def __init__(self, d_model=384, no_linear=False, only_pos=False, qk_mask=None, nhead=8, dim_feedforward=1024,
                 dropout=0.1, activation="relu", concat=False):
        super().__init__()
        self.no_linear, self.only_pos, self.qk_mask, self.concat = no_linear, only_pos, qk_mask, concat
        if concat:
            self.attn = nn.MultiheadAttention(72, nhead, vdim=d_model, dropout=dropout)
            self.newlq, self.newlk, self.outlv = nn.Linear(d_model, 72), nn.Linear(d_model, 72), nn.Linear(72, d_model)
        else:
            self.attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        if not no_linear:
            self.linear1 = nn.Linear(d_model, dim_feedforward)
            self.linear2 = nn.Linear(dim_feedforward, d_model)
            self.dropout2, self.dropout3 = nn.Dropout(dropout), nn.Dropout(dropout)
            self.norm2 = nn.LayerNorm(d_model)
            if activation not in _ACTIVATIONS:
                raise RuntimeError("activation should be relu/gelu, not %s." % activation)
            self.activation = _ACTIVATIONS[activation]

    def with_pos_embed(self, tensor, pos):
        return tensor if pos is None else tensor + pos

    def _attend(self, q, pos_q, kv, pos_kv):
        if self.concat:
            out, _ = self.attn(self.with_pos_embed(self.newlq(q), pos_q), self.with_pos_embed(self.newlk(kv), pos_kv),
                               value=kv, attn_mask=self.qk_mask)
            return q + self.outlv(self.dropout1(out))
        out, _ = self.attn(self.with_pos_embed(q, pos_q), self.with_pos_embed(kv, pos_kv), value=kv, attn_mask=self.qk_mask)
        return q + self.dropout1(out)

    def forward(self, src1_ori, pos1_ori, src2_ori, pos2_ori, attn=True):
        """src1 (queries) [B,C,N], src2 (keys / values) [B,C,M], pos*: positional embeddings of the same shapes ->
        [B,C,N].  attn=False: src2 / pos1 / pos2 are not read (None is fine)."""
        h = src1_ori.permute(2, 0, 1)                      # tokens first, as nn.MultiheadAttention wants them
        if attn:
            h = self._attend(h, pos1_ori.permute(2, 0, 1), src2_ori.permute(2, 0, 1), pos2_ori.permute(2, 0, 1))
        h = self.norm1(h)
        if not self.no_linear:
            ff = self.linear2(self.dropout2(self.activation(self.linear1(h))))
            h = self.norm2(h + self.dropout3(ff))
        return h.permute(1, 2, 0)


class TransT(nn.Module):
    """Two self blocks and two cross blocks (reference transformer.py:17-29)."""

    def __init__(self, d_model=384, concat=False):
        super().__init__()
        self.s11 = attn_module(d_model=d_model, no_linear=True, concat=concat)
        self.s12 = attn_module(d_model=d_model, no_linear=True, concat=concat)
        self.c11 = attn_module(d_model=d_model, concat=concat)
        self.c12 = attn_module(d_model=d_model, concat=concat)

    def forward(self, src1, pos1, src2, pos2, attn, need_result2=True):
        """-> (result1 [like src1], result2 [like src2]).  need_result2=False (honoured only with attn=False, where
        result1 does not depend on the src2 branch) returns (result1, None): HandTrackNet's only consumer of result2,
        ``c3`` called with attn=False, never reads it, and src2 is the (B,C,N) point feature map."""
        a = self.s11(src1, pos1, src1, pos1, attn)
        if not attn and not need_result2:
            return self.c11(a, pos1, None, pos2, attn), None
        b = self.s12(src2, pos2, src2, pos2, attn)
        return self.c11(a, pos1, b, pos2, attn), self.c12(b, pos2, a, pos1, attn)


class PositionEmbeddingSine(nn.Module):
    """sin / cos of batch-normalised coordinates at frequencies pi * 2^k (reference transformer.py:91-123):
    [B,3,N] -> [B, 6 * num_pos_feats, N]."""

    def __init__(self, num_pos_feats=64, normalize=True):
        super().__init__()
        if normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats, self.normalize = num_pos_feats, normalize

    def forward(self, coor):
        lo, hi = coor.min(), coor.max()
        unit = 2 * ((coor - lo) / (hi - lo)) - 1
        freqs = math.pi * torch.pow(2.0, torch.arange(self.num_pos_feats, dtype=torch.float, device=coor.device))
        phase = unit.unsqueeze(-1) * freqs                                   # B x 3 x N x D
        waves = torch.cat([phase.sin(), phase.cos()], dim=-1)                # B x 3 x N x 2D
        return waves.transpose(-1, -2).reshape(coor.shape[0], -1, coor.shape[-1])
