"""Hand-frame helpers: mirror of the reference's network/models/hand_utils.py for the part HandTrackNet uses
(canonicalize / decanonicalize :31-37, solve_rot_and_trans :42-66, ransac_rt :68-109, handkp2palmkp :111-124).

The rigid alignment runs on the GPU (csrc/kabsch.cu, include/pn2b200_hand.h) with its own backward instead of moving the
3x3 covariance to the CPU for torch.svd (hand_utils.py:57-61): no host synchronisation is left in HandTrackNet.forward /
compute_loss, so a whole tracking frame or training step can be captured in a CUDA graph.  The ``cpu`` arguments of the
reference signatures are accepted and ignored.
"""
import numpy as np
import torch
from torch.autograd import Function

from . import _lib


def canonicalize(data, canon_pose):  # data: [B, 3, N]
    return torch.matmul(canon_pose['rotation'].transpose(-1, -2),
                        data - canon_pose['translation']) / canon_pose['scale'][:, None, None]


def decanonicalize(data, canon_pose):  # data: [B, 3, N]
    return canon_pose['scale'][:, None, None] * torch.matmul(canon_pose['rotation'], data) + canon_pose['translation']


class _Kabsch(Function):
    @staticmethod
    def forward(ctx, x, y):
        if not (x.is_cuda and y.is_cuda):
            raise ValueError("hotrack_b200 has no CPU path")
        x = x.contiguous().float()
        y = y.contiguous().float()
        B, n, _ = y.shape
        batched = x.dim() == 3
        if batched and x.shape[0] != B:
            if x.shape[0] != 1:
                raise ValueError("x must be (num,3), (1,num,3) or (B,num,3)")
            batched, x = False, x[0].contiguous()
        R = torch.empty(B, 3, 3, dtype=torch.float32, device=y.device)
        t = torch.empty(B, 3, 1, dtype=torch.float32, device=y.device)
        aux = torch.empty(B, 12, dtype=torch.float32, device=y.device)
        _lib.call("pn2_kabsch_fwd", B, n, x.data_ptr(), 1 if batched else 0, y.data_ptr(), R.data_ptr(), t.data_ptr(),
                  aux.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(x, y, R, aux)
        ctx.batched = batched
        return R, t

    @staticmethod
    def backward(ctx, gR, gt):
        x, y, R, aux = ctx.saved_tensors
        B, n, _ = y.shape
        need_x = ctx.needs_input_grad[0] and ctx.batched
        if ctx.needs_input_grad[0] and not ctx.batched:
            raise NotImplementedError("gradient w.r.t. a template shared by the whole batch")
        gR = None if gR is None else gR.contiguous().float()
        gt = None if gt is None else gt.contiguous().float()
        gx = torch.empty_like(x) if need_x else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        if gR is None and gt is None:
            return None, None
        _lib.call("pn2_kabsch_bwd", B, n, x.data_ptr(), 1 if ctx.batched else 0, y.data_ptr(), R.data_ptr(), aux.data_ptr(),
                  0 if gR is None else gR.data_ptr(), 0 if gt is None else gt.data_ptr(),
                  0 if gx is None else gx.data_ptr(), 0 if gy is None else gy.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
        return gx, gy


def solve_rot_and_trans(x, y, cpu=True):
    """R [B,3,3], t [B,3,1] with y ~ R @ x + t (reference hand_utils.py:42-66).  x: [B,num,3] (or [num,3] shared)."""
    return _Kabsch.apply(x, y)


def ransac_rt(x, y, n=0, cpu=True):
    """Reference hand_utils.py:68-109: n == 0 is the plain least-squares fit; n in (3, 4) fits every n-subset and keeps
    the one with the smallest residual on the left-out points."""
    num = y.shape[1]
    if n == 0:
        R, t = solve_rot_and_trans(x, y)
        return R, t, None, None, None
    if n not in (3, 4):
        raise NotImplementedError
    import itertools

    index = [list(c) for c in itertools.combinations(range(num), n)]
    R_lst, t_lst, error_lst = [], [], []
    for i in index:
        R, t = solve_rot_and_trans(x[:, i, :], y[:, i, :])
        out_index = [j for j in range(num) if j not in i]
        error_lst.append((y[:, out_index, :] - torch.bmm(x[:, out_index, :], R.transpose(-1, -2))
                          - t.transpose(-1, -2)).norm(dim=-1).mean())
        R_lst.append(R)
        t_lst.append(t)
    min_ind = int(torch.stack(error_lst).argmin())
    return R_lst[min_ind], t_lst[min_ind], torch.stack(R_lst, dim=1), torch.stack(t_lst, dim=1), error_lst


def handkp2palmkp(kp):
    """[B, kp_num, 3] -> the palm keypoints [B, 6 | 14, 3] (reference hand_utils.py:111-124)."""
    # slices, not index lists: an index list becomes a host->device copy per call (a CUDA-graph capture rejects it)
    if kp.shape[1] == 21:    # joints 0, 1, 5, 9, 13, 17
        return torch.cat([kp[:, 0:2], kp[:, 5:18:4]], dim=1)
    if kp.shape[1] == 29:    # joints 0, 1, 5-7, 11-13, 17-19, 23-25
        return torch.cat([kp[:, 0:2], kp[:, 5:8], kp[:, 11:14], kp[:, 17:20], kp[:, 23:26]], dim=1)
    raise NotImplementedError
