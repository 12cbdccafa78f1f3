"""Loader for oracle/_ref/libpn2_ref.so: the REFERENCE's own CUDA kernels.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  The library is built by
`make -C oracle ref` from the unmodified sources under /root/reference (this
container only) and travels to the GPU box prebuilt.  Functions take torch CUDA
tensors and follow the reference's operator API allocation conventions
(pointnet_lib/pointnet2_utils.py), returning what its Functions return before
any sqrt / dtype cast.
"""
import ctypes
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libpn2_ref.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_SO)
    return _lib


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def furthest_point_sample(xyz, npoint, return_temp=False):
    xyz = xyz.contiguous()
    B, N, _ = xyz.shape
    out = torch.zeros(B, npoint, dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    lib().ref_furthest_point_sampling(B, N, npoint, _p(xyz), _p(temp), _p(out), _s())
    return (out, temp) if return_temp else out


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = xyz.contiguous(), new_xyz.contiguous()
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros(B, M, nsample, dtype=torch.int32, device=xyz.device)
    lib().ref_ball_query(B, N, M, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s())
    return idx


def knn(k, unknown, known):
    unknown, known = unknown.contiguous(), known.contiguous()
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = torch.zeros(B, N, k, dtype=torch.float32, device=unknown.device)
    idx = torch.zeros(B, N, k, dtype=torch.int32, device=unknown.device)
    lib().ref_knn(B, N, M, k, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return d2, idx


def three_nn(unknown, known):
    unknown, known = unknown.contiguous(), known.contiguous()
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = torch.zeros(B, N, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.zeros(B, N, 3, dtype=torch.int32, device=unknown.device)
    lib().ref_three_nn(B, N, M, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return d2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = points.contiguous(), idx.contiguous(), weight.contiguous()
    B, C, M = points.shape
    N = idx.shape[1]
    out = torch.zeros(B, C, N, dtype=torch.float32, device=points.device)
    lib().ref_three_interpolate(B, C, M, N, _p(points), _p(idx), _p(weight), _p(out), _s())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = grad_out.contiguous(), idx.contiguous(), weight.contiguous()
    B, C, N = grad_out.shape
    gp = torch.zeros(B, C, m, dtype=torch.float32, device=grad_out.device)
    lib().ref_three_interpolate_grad(B, C, N, m, _p(grad_out), _p(idx), _p(weight), _p(gp), _s())
    return gp


def group_points(points, idx):
    points, idx = points.contiguous(), idx.contiguous()
    B, C, N = points.shape
    _, S, K = idx.shape
    out = torch.zeros(B, C, S, K, dtype=torch.float32, device=points.device)
    lib().ref_group_points(B, C, N, S, K, _p(points), _p(idx), _p(out), _s())
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = grad_out.contiguous(), idx.contiguous()
    B, C, S, K = grad_out.shape
    gp = torch.zeros(B, C, n, dtype=torch.float32, device=grad_out.device)
    lib().ref_group_points_grad(B, C, n, S, K, _p(grad_out), _p(idx), _p(gp), _s())
    return gp


def gather_points(points, idx):
    points, idx = points.contiguous(), idx.contiguous()
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.zeros(B, C, M, dtype=torch.float32, device=points.device)
    lib().ref_gather_points(B, C, N, M, _p(points), _p(idx), _p(out), _s())
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = grad_out.contiguous(), idx.contiguous()
    B, C, M = grad_out.shape
    gp = torch.zeros(B, C, n, dtype=torch.float32, device=grad_out.device)
    lib().ref_gather_points_grad(B, C, n, M, _p(grad_out), _p(idx), _p(gp), _s())
    return gp
