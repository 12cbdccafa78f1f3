"""`pointnet2_cuda` as the REFERENCE's Python layer expects it, backed by the REFERENCE's own
kernels (oracle/_ref/libpn2_ref.so, compiled unmodified from /root/reference).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Stands in for the pybind module of
network/models/pointnet_lib/src/pointnet2_api.cpp:10-24, whose host wrappers no longer compile
(THC/THC.h); the ten functions keep its positional signatures and forward raw pointers to the
reference launchers on the current stream, as src/*.cpp do.  Used by oracle/ref_modules.py so the
staged reference modules (oracle/_ref/pyref) run end to end on the GPU box.
"""
import ctypes
import os

import torch

_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref", "libpn2_ref.so")
_lib = ctypes.CDLL(_SO)
_F = ctypes.c_float


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    assert t.is_cuda and t.is_contiguous(), "reference wrappers need contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _lib.ref_ball_query(b, n, m, _F(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s())
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _lib.ref_group_points(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out), _s())
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _lib.ref_group_points_grad(b, c, n, npoints, nsample, _p(grad_out), _p(idx), _p(grad_points), _s())
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _lib.ref_gather_points(b, c, n, npoints, _p(points), _p(idx), _p(out), _s())
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _lib.ref_gather_points_grad(b, c, n, npoints, _p(grad_out), _p(idx), _p(grad_points), _s())
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _lib.ref_furthest_point_sampling(b, n, m, _p(points), _p(temp), _p(idx), _s())
    return 1


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    _lib.ref_knn(b, n, m, k, _p(unknown), _p(known), _p(dist2), _p(idx), _s())


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _lib.ref_three_nn(b, n, m, _p(unknown), _p(known), _p(dist2), _p(idx), _s())


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _lib.ref_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out), _s())


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _lib.ref_three_interpolate_grad(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(grad_points), _s())
