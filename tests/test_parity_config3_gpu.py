"""BASELINE config 3 parity of the BENCHMARKED engine: the fused (tcgen05, 16-bit rows) path against the REFERENCE's own
modules (oracle/_ref/pyref: pointnet_utils.py / backbones.py, unmodified) running on the REFERENCE's own kernels
(oracle/_ref/libpn2_ref.so), strict fp32 (TF32 off), at the size the headline number is quoted on: B=32 clouds of
N=4096 points, train-mode BatchNorm.  SURVEY.md section 8d(3): indices exact, features at 16-bit tolerance (1e-2 rel),
loss curve over 50 optimiser steps tracking the fp32 run.

Tolerances (measured values in DESIGN.md section 1):
  features   norm-relative <= 1e-2 per output tensor AND element-wise |a-b| <= 4e-2 * (|b| + rms(b)) for every element
             (measured 1.5e-3 / 1.8e-3 / 3.4e-3 and 1.9e-2)
  gradients  per parameter tensor, regression loss against fixed random targets; see the test's docstring for why the
             bar is 3e-1 in the benchmarked mode and what it is compared with (gradients that are mathematically zero --
             conv biases in front of train-mode BatchNorm, SA3's last BatchNorm bias -- excepted)
  BatchNorm running statistics <= 2e-3
  loss curve over 50 Adam steps within 5e-2 of the fp32 reference's at every step
For scale the test also measures the reference against ITSELF with torch's default TF32 convolutions (what the
reference runs with on torch >= 1.12): that deviation is 20x larger than the fused engine's on the features and 4x
larger on the gradients.
"""
import types

import numpy as np
import pytest
import torch

import clouds
from oracle import ref_modules

pytestmark = pytest.mark.gpu

B, N = 32, 4096
FEAT_TOL, ELEM_TOL = 1e-2, 4e-2


@pytest.fixture(scope="module")
def ref(cuda):
    if not ref_modules.available(cuda=True):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return ref_modules.load(cuda=True)


@pytest.fixture(autouse=True)
def _fp32_exact():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _models(ref, cuda, seed=0):
    from hotrack_b200 import backbones, pointnet_utils as pu
    from hotrack_b200.handtrack_path import HandTrackPointPath, init_weights

    rpu, rbb = ref
    pu.set_engine("fused")
    try:
        ours = HandTrackPointPath(backbones.default_cfg(cuda))
    finally:
        pu.set_engine("ops")
    init_weights(ours, seed=seed)
    ns = types.SimpleNamespace(PointNet2Msg_fast=rbb.PointNet2Msg_fast,
                               PointNetSetAbstractionMsg_GivenCenterPoints=rpu.PointNetSetAbstractionMsg_GivenCenterPoints)
    theirs = HandTrackPointPath(backbones.default_cfg(cuda), ns)
    theirs.load_state_dict(ours.state_dict(), strict=True)
    return ours.to(cuda).train(), theirs.to(cuda).train()


def _inputs(cuda, seed):
    x = torch.from_numpy(clouds.ball(B, N, seed=seed)).to(cuda).transpose(1, 2).contiguous()
    k = torch.from_numpy(clouds.keypoints(B, 21, seed=seed)).to(cuda).transpose(1, 2).contiguous()
    return x, k


def _targets(outs, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn(o.shape, generator=g).to(o.device) * o.detach().std() + o.detach().mean() for o in outs]


def _loss(outs, targets):
    """Regression loss against fixed targets (the shape of HandTrackNet's keypoint loss, hand_network.py:159-221): unlike
    mean(out^2) of a BatchNorm+ReLU output, its gradient is not annihilated by the BatchNorm backward projections."""
    return sum((o - t).square().mean() for o, t in zip(outs, targets))


def _grad_table(ours, theirs_grads):
    """per-parameter norm-relative gradient error, skipping gradients that are mathematically zero on both sides (conv
    biases in front of train-mode BatchNorm; SA3's last BatchNorm bias, cancelled by FP3's BatchNorm)"""
    gmax = max(g.abs().max().item() for g in theirs_grads.values())
    out = {}
    for n1, p1 in ours.named_parameters():
        g2 = theirs_grads.get(n1)
        if g2 is None:
            continue
        if g2.abs().max().item() < 1e-5 * gmax or (n1.endswith(".bias") and "conv" in n1):
            assert p1.grad is None or p1.grad.abs().max().item() < 1e-3 * gmax, n1
            continue
        out[n1] = _rel(p1.grad, g2)
    return out


@pytest.mark.parametrize("mode", ["auto", "all"])
def test_fused_engine_matches_reference_at_config3(ref, cuda, capsys, mode):
    """mode 'auto' is the benchmarked configuration (two-plane rows in front of FP3's BatchNorm, fp16 rows behind it);
    'all' keeps two-plane rows everywhere (fp32-class forward; bench.py --precision all, 7 % slower).

    Parameter gradients: the network's max-pool / ReLU selections make the gradient a discontinuous function of the
    activations -- a forward perturbation of relative size e re-routes a fraction ~e of the selections, which moves a
    gradient that is a sum of random-sign contributions by ~sqrt(e), not e.  tools/dev/emul_grad.py shows it: the fp32
    pipeline with nothing but fp16 roundings in the forward pass behind FP3 and an EXACT fp32 backward reproduces the
    'auto' figures (4-14 %), while rounding every gradient row to bf16 with an exact forward costs < 1 %.  So 'auto' is
    held to 3e-1 per parameter tensor and to being closer to strict fp32 than the reference's own default (TF32
    convolutions) is (measured: worst 14 % / median 7 % against the reference's 59 % / 30 %); 'all' to 1.2e-1 (measured
    5.5 % / 1.8 %; features 2e-4).  The bars leave a factor ~2 over the measured worst case: the engine's own run-to-run
    noise is of the same size as these deviations."""
    from hotrack_b200 import fused

    fused.set_precise(mode)
    try:
        ours, theirs = _models(ref, cuda)
        x, k = _inputs(cuda, seed=4)
        t = theirs(x, k)
        o = ours(x, k)
    finally:
        fused.set_precise("auto")
    feat_tol, elem_tol, grad_tol = (FEAT_TOL, ELEM_TOL, 3e-1) if mode == "auto" else (5e-4, 3e-3, 1.2e-1)
    # every index tensor depends on coordinates only: exact
    for i in range(2):
        assert torch.equal(o[3][i], t[3][i])
    report, fails = [], []
    for name, a, b in zip(("src2", "f11", "f13"), o[:3], t[:3]):
        assert a.shape == b.shape and torch.isfinite(a).all()
        r = _rel(a, b)
        rms = b.square().mean().sqrt()
        elem = ((a - b).abs() / (b.abs() + rms)).max().item()
        report.append("%s: norm-rel %.2e, worst element %.2e of (|ref| + rms)" % (name, r, elem))
        if not (r < feat_tol and elem < elem_tol):
            fails.append(report[-1])
    # gradients of a regression loss against fixed random targets
    tg = _targets(t[:3], seed=1)
    _loss(t[:3], tg).backward()
    _loss(o[:3], tg).backward()
    ref_grads = {n: p.grad.clone() for n, p in theirs.named_parameters() if p.grad is not None}
    table = _grad_table(ours, ref_grads)
    worst = max(table.items(), key=lambda kv: kv[1])
    report.append("parameter gradients: worst norm-rel %.2e (%s), median %.2e" % (worst[1], worst[0], float(np.median(list(table.values())))))
    fails += ["%s grad rel %.3g" % kv for kv in table.items() if not kv[1] < grad_tol]
    for (n1, b1), (n2, b2) in zip(ours.named_buffers(), theirs.named_buffers()):
        if b1.dtype.is_floating_point:
            if not _rel(b1, b2) < 2e-3:
                fails.append("buffer %s rel %.3g" % (n1, _rel(b1, b2)))
        else:
            assert torch.equal(b1, b2), n1
    # yardstick: the reference against itself with TF32 convolutions (torch's default, i.e. how the reference runs)
    was = [bf.clone() for bf in theirs.buffers()]
    for p_ in theirs.parameters():
        p_.grad = None
    torch.backends.cudnn.allow_tf32 = True
    t32 = theirs(x, k)
    _loss(t32[:3], tg).backward()
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        for bf, w in zip(theirs.buffers(), was):
            bf.copy_(w)
    for name, a, b, c in zip(("src2", "f11", "f13"), o[:3], t[:3], t32[:3]):
        report.append("%s: reference TF32-vs-fp32 %.2e  (fused-vs-fp32 %.2e)" % (name, _rel(c.detach(), b.detach()), _rel(a.detach(), b.detach())))
        if not _rel(a.detach(), b.detach()) < _rel(c.detach(), b.detach()):
            fails.append(report[-1])
    t32_table = _grad_table(theirs, ref_grads)
    w32 = max(t32_table.items(), key=lambda kv: kv[1])
    report.append("parameter gradients, reference TF32-vs-fp32: worst %.2e (%s), median %.2e"
                  % (w32[1], w32[0], float(np.median(list(t32_table.values())))))
    if not (worst[1] < w32[1] and np.median(list(table.values())) < np.median(list(t32_table.values()))):
        fails.append(report[-1])
    with capsys.disabled():
        print("\n[config-3 parity, fused (%s) vs reference fp32, B=%d N=%d]\n  " % (mode, B, N) + "\n  ".join(report))
    assert not fails, fails


def test_loss_curve_50_steps_tracks_fp32_reference(ref, cuda, capsys):
    """50 Adam steps (lr 1e-3, weight decay 1e-4) on 5 cycling config-3 batches: the fused engine driven by TrainStep (CUDA
    graph replay, flat Adam kernel) against the reference modules + torch.optim.Adam in strict fp32."""
    from hotrack_b200.train import TrainStep

    ours, theirs = _models(ref, cuda, seed=1)
    batches = [_inputs(cuda, seed=10 + i) for i in range(5)]
    with torch.no_grad():
        was = [bf.clone() for bf in theirs.buffers()]
        tg = _targets(theirs(*batches[0])[:3], seed=2)
        for bf, w in zip(theirs.buffers(), was):
            bf.copy_(w)
    opt = torch.optim.Adam(theirs.parameters(), lr=1e-3, weight_decay=1e-4)
    ts = TrainStep(ours, lambda out: _loss(out[:3], tg), lr=1e-3, weight_decay=1e-4, graph=True)
    lr_, lo_ = [], []
    for step in range(50):
        x, k = batches[step % len(batches)]
        opt.zero_grad()
        lt = _loss(theirs(x, k)[:3], tg)
        lt.backward()
        opt.step()
        lr_.append(float(lt))
        lo_.append(float(ts(x, k)))
    lr_, lo_ = np.array(lr_), np.array(lo_)
    dev = np.abs(lo_ - lr_) / lr_
    with capsys.disabled():
        print("\n[loss curve, 50 steps] reference %.4f -> %.4f, fused %.4f -> %.4f, max relative gap %.2e (step %d)"
              % (lr_[0], lr_[-1], lo_[0], lo_[-1], dev.max(), int(dev.argmax())))
    assert np.isfinite(lo_).all()
    assert lr_[-1] < 0.9 * lr_[0], "the reference run must actually learn for the comparison to mean something"
    assert dev.max() < 5e-2, dev   # measured 1.7e-2 .. 2.1e-2 (atomics make runs differ), growing with the step count
    assert ts.opt.t == 50
    pr = torch.cat([p.detach().flatten() for p in theirs.parameters()])
    po = torch.cat([p.detach().flatten() for p in ours.parameters()])
    assert _rel(po, pr) < 1.5e-1, _rel(po, pr)  # measured 8e-2: 50 updates of lr-sized (sign-like) Adam steps from gradients that differ by ~7 %
