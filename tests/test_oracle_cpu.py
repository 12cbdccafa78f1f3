"""CPU suite: pins the oracle (oracle/pn2_oracle.c) to the golden vectors captured from the
REFERENCE's own kernels on a B200 (tests/golden/*.npz, made by tools/make_golden.py), and checks
domain properties + the edge cases the reference semantics define.

No GPU needed.  The oracle is the checker for the `-m gpu` parity tests, so it is itself checked here.
"""
import glob
import os

import numpy as np
import pytest

import clouds
from oracle import pn2_oracle as orc

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    xyz, new_xyz, kps = g["xyz"], g["new_xyz"], g["kps"]
    npoint, nsample, k = int(g["npoint"]), int(g["nsample"]), int(g["k"])
    idx, temp = orc.furthest_point_sample(xyz, npoint, return_temp=True)
    np.testing.assert_array_equal(idx, g["fps_idx"])
    np.testing.assert_array_equal(temp, g["fps_temp"])
    np.testing.assert_array_equal(orc.ball_query(float(g["radius"]), nsample, xyz, new_xyz), g["ball_idx"])
    d2, kidx = orc.knn(k, kps, xyz)
    np.testing.assert_array_equal(kidx, g["knn_idx"])
    np.testing.assert_array_equal(d2, g["knn_d2"])
    d2, nidx = orc.three_nn(xyz, new_xyz)
    np.testing.assert_array_equal(nidx, g["nn_idx"])
    np.testing.assert_array_equal(d2, g["nn_d2"])
    np.testing.assert_array_equal(orc.three_interpolate(g["feats"], g["nn_idx"], g["weight"]), g["interp"])
    xt = np.ascontiguousarray(xyz.transpose(0, 2, 1))
    np.testing.assert_array_equal(orc.group_points(xt, g["ball_idx"]), g["grouped_xyz"])
    np.testing.assert_array_equal(orc.gather_points(xt, g["fps_idx"]), g["gathered_xyz"])


# ---------------------------------------------------------------- FPS -------
def _fps_numpy(xyz, m):
    """Independent restatement: sequential scan + explicit tie rule
    (max distance, then min (bitrev(k mod bs), k div bs)); SURVEY.md section 2b."""
    B, N, _ = xyz.shape
    bs = orc.lib().pn2o_opt_n_threads(N)
    lg = bs.bit_length() - 1
    k = np.arange(N)
    rev = np.array([int(format(v, "0%db" % lg)[::-1], 2) if lg else 0 for v in (k % bs)])
    prio = rev * ((N + bs - 1) // bs) + k // bs
    out = np.zeros((B, m), np.int32)
    for b in range(B):
        temp = np.full(N, 1e10, np.float32)
        old = 0
        for j in range(1, m):
            d = (xyz[b] - xyz[b, old]).astype(np.float32)
            dx, dy, dz = d[:, 0].astype(np.float64), d[:, 1].astype(np.float64), d[:, 2].astype(np.float64)
            t = (dy * dy).astype(np.float32).astype(np.float64)
            t = (dx * dx + t).astype(np.float32).astype(np.float64)  # fma: one rounding
            d2 = (dz * dz + t).astype(np.float32)
            temp = np.minimum(d2, temp)
            best = temp.max()
            cand = np.nonzero(temp == best)[0]
            old = int(cand[np.argmin(prio[cand])])
            out[b, j] = old
    return out


@pytest.mark.parametrize("kind,B,N,M", [("ball", 2, 1024, 64), ("lattice", 2, 1000, 96), ("lattice", 1, 300, 300),
                                         ("duplicates", 1, 2560, 40), ("ball", 1, 5, 5), ("lattice", 1, 33, 40)])
def test_fps_tie_rule_restatement(kind, B, N, M):
    xyz = clouds.make(kind, B, N, seed=N)
    np.testing.assert_array_equal(orc.furthest_point_sample(xyz, M), _fps_numpy(xyz, M))


def test_fps_properties():
    xyz = clouds.ball(3, 2048, seed=3)
    idx, temp = orc.furthest_point_sample(xyz, 256, return_temp=True)
    assert (idx[:, 0] == 0).all()
    for b in range(3):
        assert len(set(idx[b].tolist())) == 256  # distinct points on a generic cloud
        # running min distance of the chosen points is non-increasing (greedy farthest-first)
        sel = xyz[b][idx[b]]
        dmin = [np.min(np.sum((sel[:j] - sel[j]) ** 2, -1)) for j in range(1, 256)]
        assert all(dmin[i] >= dmin[i + 1] - 1e-7 for i in range(len(dmin) - 1))
        assert temp[b][idx[b][:-1]].max() == 0.0  # sampled points are at distance 0 from the set
    assert orc.furthest_point_sample(xyz, 0).shape == (3, 0)


def test_fps_block_size_rule():
    f = orc.lib().pn2o_opt_n_threads
    assert [f(n) for n in (1, 2, 3, 255, 256, 1000, 1024, 2560, 4096, 8192)] == [1, 2, 2, 128, 256, 512, 1024, 1024,
                                                                                   1024, 1024]


# ---------------------------------------------------------- ball query -------
def test_ball_query_semantics():
    xyz = clouds.lattice(2, 2048, seed=5)
    centres = xyz[:, :64].copy()
    centres[:, -1] += 100.0  # empty ball
    r, K = 0.125, 16  # lattice step 1/16: points exactly ON the radius are excluded (strict <)
    idx = orc.ball_query(r, K, xyz, centres)
    r2 = np.float32(r) * np.float32(r)
    for b in range(2):
        for s in range(64):
            d = (centres[b, s] - xyz[b]).astype(np.float32)
            d2 = (d.astype(np.float64) ** 2).sum(-1)  # exact on the lattice
            hits = np.nonzero(d2 < r2)[0]
            want = np.zeros(K, np.int32)
            if len(hits):
                want[:] = hits[0]
                want[: min(K, len(hits))] = hits[:K]
            np.testing.assert_array_equal(idx[b, s], want)
    assert (idx[:, -1] == 0).all()


# ------------------------------------------------------- kNN / three_nn ------
@pytest.mark.parametrize("kind", ["ball", "lattice", "duplicates"])
def test_knn_is_stable_sorted_prefix(kind):
    known = clouds.make(kind, 2, 700, seed=11)
    unknown = clouds.lattice(2, 9, seed=12) if kind == "lattice" else clouds.keypoints(2, 9, seed=12)
    k = 40
    d2, idx = orc.knn(k, unknown, known)
    assert (np.diff(d2, axis=-1) >= 0).all()  # sortedness
    for b in range(2):
        for q in range(9):
            d = (unknown[b, q] - known[b]).astype(np.float32)
            full = (d[:, 1] * d[:, 1]).astype(np.float32)
            full = np.float32(d[:, 0].astype(np.float64) * d[:, 0] + full)
            full = np.float32(d[:, 2].astype(np.float64) * d[:, 2] + full.astype(np.float64))
            order = np.lexsort((np.arange(700), full))[:k]  # (distance, index) lexicographic
            np.testing.assert_array_equal(idx[b, q], order)
            np.testing.assert_array_equal(d2[b, q], full[order])
    # three_nn is knn with k = 3
    d3, i3 = orc.three_nn(unknown, known)
    np.testing.assert_array_equal(i3, idx[..., :3])
    np.testing.assert_array_equal(d3, d2[..., :3])


def test_knn_fewer_points_than_k():
    known = clouds.ball(1, 2, seed=1)
    unknown = clouds.keypoints(1, 4, seed=1)
    d2, idx = orc.knn(5, unknown, known)
    assert np.isinf(d2[..., 2:]).all() and (idx[..., 2:] == 0).all()
    d3, i3 = orc.three_nn(unknown, known)
    assert np.isinf(d3[..., 2]).all() and (i3[..., 2] == 0).all()


# ----------------------------------------- interpolate / group / gather ------
def test_interpolate_linearity_and_grad_adjoint():
    rng = np.random.RandomState(0)
    B, C, m, n = 2, 6, 50, 200
    idx = rng.randint(0, m, size=(B, n, 3)).astype(np.int32)
    w = rng.rand(B, n, 3).astype(np.float32)
    p = rng.randn(B, C, m).astype(np.float32)
    g = rng.randn(B, C, n).astype(np.float32)
    out = orc.three_interpolate(p, idx, w)
    # exact definition with the reference's rounding sequence: fma(w2,p2, fma(w0,p0, rn(w1*p1)))
    for b in range(B):
        p0, p1, p2 = (p[b][:, idx[b, :, j]] for j in range(3))
        t = (w[b, :, 1] * p1).astype(np.float32)
        t = (w[b, :, 0].astype(np.float64) * p0 + t).astype(np.float32)
        want = (w[b, :, 2].astype(np.float64) * p2 + t).astype(np.float32)
        np.testing.assert_array_equal(out[b], want)
    # <interp(p), g> == <p, interp_grad(g)> (adjoint), to fp32 accumulation error
    gp = orc.three_interpolate_grad(g, idx, w, m)
    lhs = float((out.astype(np.float64) * g).sum())
    rhs = float((p.astype(np.float64) * gp).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))


def test_group_gather_are_exact_copies_and_grads_scatter():
    rng = np.random.RandomState(1)
    B, C, N, S, K = 2, 5, 40, 7, 6
    p = rng.randn(B, C, N).astype(np.float32)
    idx = rng.randint(0, N, size=(B, S, K)).astype(np.int32)
    out = orc.group_points(p, idx)
    for b in range(B):
        np.testing.assert_array_equal(out[b], p[b][:, idx[b]])
    g = rng.randn(B, C, S, K).astype(np.float32)
    gp = orc.group_points_grad(g, idx, N)
    want = np.zeros((B, C, N), np.float64)
    for b in range(B):
        for s in range(S):
            for k in range(K):
                want[b, :, idx[b, s, k]] += g[b, :, s, k]
    np.testing.assert_allclose(gp, want, rtol=1e-6, atol=1e-6)
    gi = idx[:, :, 0].copy()
    np.testing.assert_array_equal(orc.gather_points(p, gi), np.stack([p[b][:, gi[b]] for b in range(B)]))
    gg = orc.gather_points_grad(g[..., 0].copy(), gi, N)
    want = np.zeros((B, C, N), np.float64)
    for b in range(B):
        for s in range(S):
            want[b, :, gi[b, s]] += g[b, :, s, 0]
    np.testing.assert_allclose(gg, want, rtol=1e-6, atol=1e-6)
