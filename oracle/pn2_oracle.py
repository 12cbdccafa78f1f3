"""numpy front-end of oracle/pn2_oracle.c (CPU restatement of the reference kernels).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Function names and argument
meaning follow the reference's operator API (pointnet_lib/pointnet2_utils.py);
allocation conventions are the reference's (temp = 1e10, zero-filled ball-query
and grad outputs: pointnet2_utils.py:27-28,71,186,232,262).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpn2_oracle.so")
_SRC = os.path.join(_HERE, "pn2_oracle.c")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def furthest_point_sample(xyz, npoint, return_temp=False):
    xyz = _f(xyz)
    B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, dtype=np.float32)
    idx = np.zeros((B, npoint), dtype=np.int32)
    lib().pn2o_furthest_point_sampling(B, N, npoint, _p(xyz), _p(temp), _p(idx))
    return (idx, temp) if return_temp else idx


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), dtype=np.int32)
    lib().pn2o_ball_query(B, N, M, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx))
    return idx


def knn(k, unknown, known):
    """Returns (dist2, idx): squared distances, as the kernel writes them."""
    unknown, known = _f(unknown), _f(known)
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = np.zeros((B, N, k), dtype=np.float32)
    idx = np.zeros((B, N, k), dtype=np.int32)
    lib().pn2o_knn(B, N, M, k, _p(unknown), _p(known), _p(d2), _p(idx))
    return d2, idx


def three_nn(unknown, known):
    unknown, known = _f(unknown), _f(known)
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = np.zeros((B, N, 3), dtype=np.float32)
    idx = np.zeros((B, N, 3), dtype=np.int32)
    lib().pn2o_three_nn(B, N, M, _p(unknown), _p(known), _p(d2), _p(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f(points), _i(idx), _f(weight)
    B, C, M = points.shape
    N = idx.shape[1]
    out = np.zeros((B, C, N), dtype=np.float32)
    lib().pn2o_three_interpolate(B, C, M, N, _p(points), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f(grad_out), _i(idx), _f(weight)
    B, C, N = grad_out.shape
    gp = np.zeros((B, C, m), dtype=np.float32)
    lib().pn2o_three_interpolate_grad(B, C, N, m, _p(grad_out), _p(idx), _p(weight), _p(gp))
    return gp


def group_points(points, idx):
    points, idx = _f(points), _i(idx)
    B, C, N = points.shape
    _, S, K = idx.shape
    out = np.zeros((B, C, S, K), dtype=np.float32)
    lib().pn2o_group_points(B, C, N, S, K, _p(points), _p(idx), _p(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f(grad_out), _i(idx)
    B, C, S, K = grad_out.shape
    gp = np.zeros((B, C, n), dtype=np.float32)
    lib().pn2o_group_points_grad(B, C, n, S, K, _p(grad_out), _p(idx), _p(gp))
    return gp


def gather_points(points, idx):
    points, idx = _f(points), _i(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.zeros((B, C, M), dtype=np.float32)
    lib().pn2o_gather_points(B, C, N, M, _p(points), _p(idx), _p(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f(grad_out), _i(idx)
    B, C, M = grad_out.shape
    gp = np.zeros((B, C, n), dtype=np.float32)
    lib().pn2o_gather_points_grad(B, C, n, M, _p(grad_out), _p(idx), _p(gp))
    return gp


def opt_n_threads(n):
    return lib().pn2o_opt_n_threads(int(n))
