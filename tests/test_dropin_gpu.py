"""The drop-in proof: the REFERENCE's UNMODIFIED Python layer -- pointnet_lib/pointnet2_utils.py (its seven
autograd.Functions, allocating with torch.cuda.IntTensor / FloatTensor exactly as it does), pointnet_utils.py,
backbones.py and the whole HandTrackNet of hand_network.py -- imported twice from oracle/_ref/pyref: once with
``import pointnet2_cuda`` resolving to the reference's own kernels (oracle/refshim -> libpn2_ref.so) and once resolving to
hotrack_b200/dropin/pointnet2_cuda.py (libpn2b200.so).  Same weights, same inputs, strict fp32:
every index tensor identical, every float output within 1e-5 (north_star), BASELINE config 2 (B=8, N=2048).
"""
import numpy as np
import pytest
import torch

import clouds
from oracle import ref_modules

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def both(cuda):
    if not ref_modules.available(cuda=True, full=True):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return ref_modules.load_full("ref"), ref_modules.load_full("ours")


@pytest.fixture(autouse=True)
def _fp32_exact():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_reference_operator_api_on_this_backend(both, cuda):
    """pointnet_lib/pointnet2_utils.py's Functions, unmodified, on both backends: exact integers, exact forward floats."""
    (rpu, _, _), (opu, _, _) = both
    rf, of = rpu.futils, opu.futils
    assert rf is not of and rf.pointnet2 is not of.pointnet2
    B, N = 4, 2048
    xyz = torch.from_numpy(clouds.duplicates(B, N, seed=3)).to(cuda)
    kp = torch.from_numpy(clouds.keypoints(B, 21, seed=3)).to(cuda)
    ir, io = rf.furthest_point_sample(xyz, 256), of.furthest_point_sample(xyz, 256)
    assert ir.dtype == io.dtype == torch.int32 and torch.equal(ir, io)
    new_xyz = torch.gather(xyz, 1, ir.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    br, bo = rf.ball_query(0.1, 32, xyz, new_xyz), of.ball_query(0.1, 32, xyz, new_xyz)
    assert torch.equal(br, bo)
    (dr, kr), (do, ko) = rf.knn(64, kp, xyz), of.knn(64, kp, xyz)
    assert torch.equal(kr, ko) and torch.equal(dr, do)
    (dr, tr), (do, to) = rf.three_nn(xyz, new_xyz), of.three_nn(xyz, new_xyz)
    assert torch.equal(tr, to) and torch.equal(dr, do)
    feats = torch.randn(B, 16, 256, device=cuda)
    w = torch.rand(B, N, 3, device=cuda)
    w = w / w.sum(-1, keepdim=True)
    fr = feats.clone().requires_grad_(True)
    fo = feats.clone().requires_grad_(True)
    yr, yo = rf.three_interpolate(fr, tr, w), of.three_interpolate(fo, to, w)
    assert torch.equal(yr, yo)
    g = torch.randn_like(yr)
    yr.backward(g)
    yo.backward(g)
    assert _rel(fo.grad, fr.grad) < 1e-5  # atomics: summation order
    pts = torch.randn(B, 16, N, device=cuda)
    pr = pts.clone().requires_grad_(True)
    po = pts.clone().requires_grad_(True)
    gr, go = rf.grouping_operation(pr, br), of.grouping_operation(po, bo)
    assert torch.equal(gr, go)
    g = torch.randn_like(gr)
    gr.backward(g)
    go.backward(g)
    assert _rel(po.grad, pr.grad) < 1e-5
    gr, go = rf.gather_operation(pts, ir), of.gather_operation(pts, io)
    assert torch.equal(gr, go)


@pytest.mark.parametrize("train", [False, True])
def test_reference_handtracknet_unmodified_on_this_backend(both, cuda, train):
    """HandTrackNet(cfg) of the reference's hand_network.py:45-157 (handframe 'camera'), config 2: same pred_kp on both
    backends; in training mode also the same loss and parameter gradients through the reference's own backward."""
    (rpu, rbb, rhn), (opu, obb, ohn) = both
    B, N = 8, 2048
    torch.manual_seed(0)
    cfg = ref_modules.handtracknet_cfg(cuda, "camera")
    net_r = rhn.HandTrackNet(cfg).to(cuda)
    net_o = ohn.HandTrackNet(cfg).to(cuda)
    net_o.load_state_dict(net_r.state_dict(), strict=True)
    net_r.train(train)
    net_o.train(train)
    # metres, as the data loader delivers them (the network divides by scale 0.2: hand_network.py:99,118-119)
    data = {"hand_points": torch.from_numpy(clouds.ball(B, N, seed=2)) * 0.2,
            "jittered_hand_kp": torch.from_numpy(clouds.keypoints(B, 21, seed=2)) * 0.2,
            "gt_hand_kp": torch.from_numpy(clouds.keypoints(B, 21, seed=3)) * 0.2,
            "gt_hand_pose": {"palm_template": torch.from_numpy(clouds.keypoints(B, 6, seed=4)) * 0.2}}
    flags = {"track_flag": False, "IKNet_flag": train is False}
    outs = []
    for net in (net_r, net_o):
        torch.manual_seed(1)  # the dropout masks of the (live) feed-forward blocks in training mode
        if train:
            ret = net(data, flags)
            loss = (ret["pred_kp"] - data["gt_hand_kp"].to(cuda)).abs().mean()
            loss.backward()
        else:
            with torch.no_grad():
                ret = net(data, flags)
        outs.append(ret)
    r, o = outs
    assert r["pred_kp"].shape == (B, 21, 3)
    assert _rel(o["pred_kp"], r["pred_kp"]) < 1e-5
    assert _rel(o["pred_kp_handframe"], r["pred_kp_handframe"]) < 1e-5
    if not train:
        assert torch.equal(o["pred_kp_vis_mask"], r["pred_kp_vis_mask"])  # K=4 kNN distances -> visibility (hand_network.py:149-155)
        return
    gmax = max(p.grad.abs().max().item() for p in net_r.parameters() if p.grad is not None)
    for (n1, p1), (n2, p2) in zip(net_o.named_parameters(), net_r.named_parameters()):
        assert n1 == n2 and (p1.grad is None) == (p2.grad is None), n1
        if p2.grad is None or p2.grad.abs().max().item() < 1e-5 * gmax:
            continue
        if n1.endswith(".bias") and "conv" in n1 and "final_mlp" not in n1 and ".linear" not in n1:
            continue  # in front of train-mode BatchNorm: rounding noise around an exact zero on both sides
        assert _rel(p1.grad, p2.grad) < 1e-3, (n1, _rel(p1.grad, p2.grad))
