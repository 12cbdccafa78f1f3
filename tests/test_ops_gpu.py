"""Parity of the sm_100a kernels (through the C ABI) against the CPU oracle and,
when oracle/_ref was built, against the reference's own kernels on the same GPU.

Integer outputs (FPS / ball_query / kNN / three_nn indices) must be bit-exact;
forward float outputs are bit-exact too (same rounding sequence); atomics-based
gradients are compared to 1e-5 relative (summation order differs, as it does
between two runs of the reference itself).
"""
import numpy as np
import pytest
import torch

import clouds
from oracle import pn2_oracle as orc
from oracle import ref_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from hotrack_b200 import pointnet2_utils as futils

    return futils


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


FPS_CASES = [
    # kind, B, N, npoint
    ("ball", 2, 1024, 256),      # BASELINE config 1
    ("ball", 3, 4096, 256),      # SA1 shape
    ("shell", 2, 256, 128),      # SA2 shape
    ("ball", 2, 2048, 256),      # config 2
    ("ball", 1, 8192, 256),      # config 5
    ("lattice", 2, 4096, 256),   # exact ties everywhere
    ("lattice", 2, 1000, 128),   # reference block 512, ragged second row
    ("duplicates", 2, 2560, 512),  # data-loader shape, block 1024, 3 rows (ragged)
    ("lattice", 2, 300, 300),    # npoint == N with duplicates: zeros tail
    ("ball", 1, 21, 8),
    ("ball", 2, 5, 5),
    ("ball", 1, 1, 1),
    ("lattice", 1, 33, 40),      # npoint > N
    ("ball", 1, 10000, 64),      # streaming path (needs temp)
]


@pytest.mark.parametrize("kind,B,N,M", FPS_CASES)
def test_fps_matches_oracle_and_reference(ops, cuda, kind, B, N, M):
    from hotrack_b200 import pointnet2_cuda as pc

    xyz = clouds.make(kind, B, N, seed=N + M)
    want, want_temp = orc.furthest_point_sample(xyz, M, return_temp=True)
    x = _t(xyz, cuda)
    temp = torch.full((B, N), 1e10, device=cuda)
    got = torch.zeros(B, M, dtype=torch.int32, device=cuda)
    pc.furthest_point_sampling_wrapper(B, N, M, x, temp, got)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(temp.cpu().numpy(), want_temp)  # scratch contract kept too
    if N <= 8192:
        got2 = ops.furthest_point_sample(x, M)  # temp-less path used by the operator API
        np.testing.assert_array_equal(got2.cpu().numpy(), want)
    if ref_lib.available():
        np.testing.assert_array_equal(ref_lib.furthest_point_sample(x, M).cpu().numpy(), want)


BALL_CASES = [
    # kind, B, N, S, radius, nsample
    ("ball", 2, 1024, 256, 0.1, 32),    # config 1
    ("ball", 2, 4096, 256, 0.1, 32),    # SA1
    ("shell", 2, 256, 128, 0.2, 32),    # SA2
    ("lattice", 2, 2048, 128, 0.125, 32),  # points exactly ON the radius (strict <)
    ("lattice", 2, 2048, 64, 0.25, 64),
    ("ball", 3, 1023, 77, 0.15, 16),    # N % 4 != 0: unaligned clouds, no bulk copy
    ("ball", 1, 4097, 50, 0.1, 32),     # ragged last tile
    ("ball", 2, 5000, 100, 0.05, 8),
    ("ball", 2, 3, 2, 0.5, 4),
    ("duplicates", 2, 2048, 128, 0.2, 32),
]


@pytest.mark.parametrize("kind,B,N,S,radius,nsample", BALL_CASES)
def test_ball_query(ops, cuda, kind, B, N, S, radius, nsample):
    xyz = clouds.make(kind, B, N, seed=N + S)
    centres = xyz[:, np.random.RandomState(S).permutation(N)[:S]].copy()
    centres[:, -1] += 10.0  # one centre with an empty ball: row must stay zero
    want = orc.ball_query(radius, nsample, xyz, centres)
    got = ops.ball_query(radius, nsample, _t(xyz, cuda), _t(centres, cuda))
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert (want[:, -1] == 0).all()
    if ref_lib.available():
        ref = ref_lib.ball_query(radius, nsample, _t(xyz, cuda), _t(centres, cuda))
        np.testing.assert_array_equal(ref.cpu().numpy(), want)


KNN_CASES = [
    # kind, B, n, m, k
    ("ball", 2, 21, 4096, 16),   # HandTrackNet q1
    ("ball", 2, 21, 4096, 64),
    ("ball", 2, 21, 2048, 4),    # visibility test
    ("lattice", 2, 21, 4096, 64),  # ties resolved by index
    ("duplicates", 2, 37, 1500, 33),
    ("ball", 1, 300, 1025, 200),  # reference maximum k
    ("ball", 2, 5, 10, 16),      # m < k: (+inf, 0) tail
    ("ball", 1, 9, 3001, 128),
]


@pytest.mark.parametrize("kind,B,n,m,k", KNN_CASES)
def test_knn(ops, cuda, kind, B, n, m, k):
    from hotrack_b200 import pointnet2_cuda as pc

    known = clouds.make(kind, B, m, seed=m + k)
    unknown = clouds.keypoints(B, n, seed=k) if kind != "lattice" else clouds.lattice(B, n, seed=k + 1)
    want_d2, want_idx = orc.knn(k, unknown, known)
    u, kn = _t(unknown, cuda), _t(known, cuda)
    d2 = torch.empty(B, n, k, device=cuda)
    idx = torch.empty(B, n, k, dtype=torch.int32, device=cuda)
    pc.knn_wrapper(B, n, m, k, u, kn, d2, idx)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(d2.cpu().numpy(), want_d2)
    dist, idx2 = ops.knn(k, u, kn)
    np.testing.assert_array_equal(idx2.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(dist.cpu().numpy(), np.sqrt(want_d2))
    if ref_lib.available():
        rd2, ridx = ref_lib.knn(k, u, kn)
        np.testing.assert_array_equal(ridx.cpu().numpy(), want_idx)
        np.testing.assert_array_equal(rd2.cpu().numpy(), want_d2)


NN3_CASES = [("ball", 2, 4096, 256), ("shell", 2, 256, 128), ("lattice", 2, 1000, 300), ("ball", 2, 100, 2),
             ("ball", 1, 50, 1), ("duplicates", 2, 777, 2500)]


@pytest.mark.parametrize("kind,B,n,m", NN3_CASES)
def test_three_nn(ops, cuda, kind, B, n, m):
    unknown = clouds.make(kind, B, n, seed=n)
    known = clouds.make(kind, B, m, seed=m + 1)
    want_d2, want_idx = orc.three_nn(unknown, known)
    dist, idx = ops.three_nn(_t(unknown, cuda), _t(known, cuda))
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(dist.cpu().numpy(), np.sqrt(want_d2))
    if ref_lib.available():
        rd2, ridx = ref_lib.three_nn(_t(unknown, cuda), _t(known, cuda))
        np.testing.assert_array_equal(ridx.cpu().numpy(), want_idx)
        np.testing.assert_array_equal(rd2.cpu().numpy(), want_d2)


INTERP_CASES = [(2, 128, 256, 4096), (2, 256, 128, 256), (1, 7, 5, 33), (2, 40, 3000, 1000), (1, 3, 9000, 100)]


@pytest.mark.parametrize("B,C,m,n", INTERP_CASES)
def test_three_interpolate_fwd_bwd(ops, cuda, B, C, m, n):
    rng = np.random.RandomState(C + m)
    pts = rng.randn(B, C, m).astype(np.float32)
    idx = rng.randint(0, m, size=(B, n, 3)).astype(np.int32)
    w = rng.rand(B, n, 3).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    want = orc.three_interpolate(pts, idx, w)
    p = _t(pts, cuda).requires_grad_(True)
    out = ops.three_interpolate(p, _t(idx, cuda), _t(w, cuda))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), want)  # same rounding sequence: bit-exact
    g = rng.randn(B, C, n).astype(np.float32)
    out.backward(_t(g, cuda))
    want_g = orc.three_interpolate_grad(g, idx, w, m)
    np.testing.assert_allclose(p.grad.cpu().numpy(), want_g, rtol=1e-5, atol=1e-5)
    if ref_lib.available():
        ref = ref_lib.three_interpolate(_t(pts, cuda), _t(idx, cuda), _t(w, cuda))
        np.testing.assert_array_equal(ref.cpu().numpy(), want)
        rg = ref_lib.three_interpolate_grad(_t(g, cuda), _t(idx, cuda), _t(w, cuda), m)
        np.testing.assert_allclose(rg.cpu().numpy(), want_g, rtol=1e-5, atol=1e-5)


GROUP_CASES = [(2, 3, 1024, 256, 32), (2, 64, 256, 128, 32), (1, 5, 17, 3, 2), (2, 384, 4096, 21, 64)]


@pytest.mark.parametrize("B,C,N,S,K", GROUP_CASES)
def test_group_and_gather_fwd_bwd(ops, cuda, B, C, N, S, K):
    rng = np.random.RandomState(C + N)
    pts = rng.randn(B, C, N).astype(np.float32)
    idx = rng.randint(0, N, size=(B, S, K)).astype(np.int32)
    p = _t(pts, cuda).requires_grad_(True)
    out = ops.grouping_operation(p, _t(idx, cuda))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), orc.group_points(pts, idx))
    g = rng.randn(B, C, S, K).astype(np.float32)
    out.backward(_t(g, cuda))
    np.testing.assert_allclose(p.grad.cpu().numpy(), orc.group_points_grad(g, idx, N), rtol=1e-5, atol=1e-5)

    gidx = idx[:, :, 0].copy()
    p2 = _t(pts, cuda).requires_grad_(True)
    out2 = ops.gather_operation(p2, _t(gidx, cuda))
    np.testing.assert_array_equal(out2.detach().cpu().numpy(), orc.gather_points(pts, gidx))
    g2 = rng.randn(B, C, S).astype(np.float32)
    out2.backward(_t(g2, cuda))
    np.testing.assert_allclose(p2.grad.cpu().numpy(), orc.gather_points_grad(g2, gidx, N), rtol=1e-5, atol=1e-5)
    if ref_lib.available():
        np.testing.assert_array_equal(ref_lib.group_points(_t(pts, cuda), _t(idx, cuda)).cpu().numpy(),
                                      orc.group_points(pts, idx))


def test_config1_single_sa_layer_indices(ops, cuda):
    """BASELINE.json config 1: B=2, N=1024, npoint=256, nsample=32, C=3 -- FPS + ball_query + group."""
    xyz = clouds.ball(2, 1024, seed=1)
    x = _t(xyz, cuda)
    fps = ops.furthest_point_sample(x, 256)
    np.testing.assert_array_equal(fps.cpu().numpy(), orc.furthest_point_sample(xyz, 256))
    new_xyz = ops.gather_operation(x.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    want_new = np.stack([xyz[b][fps[b].cpu().numpy()] for b in range(2)])
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), want_new)
    idx = ops.ball_query(0.1, 32, x, new_xyz)
    np.testing.assert_array_equal(idx.cpu().numpy(), orc.ball_query(0.1, 32, xyz, want_new))
    grouped = ops.grouping_operation(x.transpose(1, 2).contiguous(), idx)
    np.testing.assert_array_equal(grouped.cpu().numpy(),
                                  orc.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx.cpu().numpy()))


def test_bad_arguments_raise(ops, cuda):
    from hotrack_b200 import _lib, pointnet2_cuda as pc

    x = torch.zeros(1, 8, 3, device=cuda)
    with pytest.raises(_lib.Pn2Error):
        pc.knn_wrapper(1, 8, 8, 5000, x, x, torch.zeros(1, 8, 5000, device=cuda),
                       torch.zeros(1, 8, 5000, dtype=torch.int32, device=cuda))
    with pytest.raises(ValueError):
        pc.three_nn_wrapper(1, 8, 8, x.cpu(), x, x, x)


def test_data_loader_fps_matches_reference_semantics(cuda):
    """datasets/data_utils.farthest_point_sample (B=1, <= 5*npoint points, npoint=512) and the
    batched variant; checked against the oracle on the points the function actually samples from."""
    from hotrack_b200 import data_utils

    rng = np.random.RandomState(0)
    small = clouds.ball(1, 2000, seed=31)[0]          # <= 5*512: used as is
    big = clouds.shell(1, 9000, seed=32)[0]           # > 5*512: random 2560-subset first
    np.random.seed(5)
    got = data_utils.farthest_point_sample(small, 512, cuda)
    np.testing.assert_array_equal(got, orc.furthest_point_sample(small[None], 512)[0])
    np.random.seed(5)
    got = data_utils.farthest_point_sample(big, 512, cuda)
    np.random.seed(5)
    sub = np.random.permutation(len(big))[:2560]
    np.testing.assert_array_equal(got, sub[orc.furthest_point_sample(big[sub][None], 512)[0]])
    # batched: ragged clouds in one launch == one call per cloud
    cl = [clouds.ball(1, n, seed=40 + n)[0] for n in (700, 2560, 1500)]
    batch = data_utils.sample_batch(cl, 512, cuda)
    for c, b in zip(cl, batch):
        np.testing.assert_array_equal(b, orc.furthest_point_sample(c[None], 512)[0])
