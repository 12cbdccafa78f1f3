"""Drop-in for the reference's ``pointnet2_cuda`` extension module.

Same ten functions, same positional arguments and return values as the pybind
module the reference builds from network/models/pointnet_lib/src/pointnet2_api.cpp:10-24,
so ``import pointnet2_cuda as pointnet2`` (pointnet_lib/pointnet2_utils.py:7)
resolves to this file when ``hotrack_b200/dropin`` is on ``sys.path``.

Each wrapper forwards raw device pointers to the C ABI in include/pn2b200.h on
the current CUDA stream, exactly as the reference wrappers do
(src/sampling.cpp:38-49, src/ball_query.cpp:14-25, src/interpolate.cpp:14-69,
src/group_points.cpp).  Differences: arguments are validated (the reference
checks only ball_query's inputs) and a failed launch raises instead of exit(-1).
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (hotrack_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t.data_ptr()


_f32, _i32 = torch.float32, torch.int32


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _lib.call("pn2_ball_query", b, n, m, float(radius), nsample, _ptr(new_xyz, _f32, "new_xyz"),
              _ptr(xyz, _f32, "xyz"), _ptr(idx, _i32, "idx"), _stream())
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _lib.call("pn2_group_points", b, c, n, npoints, nsample, _ptr(points, _f32, "points"),
              _ptr(idx, _i32, "idx"), _ptr(out, _f32, "out"), _stream())
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _lib.call("pn2_group_points_grad", b, c, n, npoints, nsample, _ptr(grad_out, _f32, "grad_out"),
              _ptr(idx, _i32, "idx"), _ptr(grad_points, _f32, "grad_points"), _stream())
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _lib.call("pn2_gather_points", b, c, n, npoints, _ptr(points, _f32, "points"), _ptr(idx, _i32, "idx"),
              _ptr(out, _f32, "out"), _stream())
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _lib.call("pn2_gather_points_grad", b, c, n, npoints, _ptr(grad_out, _f32, "grad_out"),
              _ptr(idx, _i32, "idx"), _ptr(grad_points, _f32, "grad_points"), _stream())
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _lib.call("pn2_furthest_point_sampling", b, n, m, _ptr(points, _f32, "points"),
              None if temp is None else _ptr(temp, _f32, "temp"), _ptr(idx, _i32, "idx"), _stream())
    return 1


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    _lib.call("pn2_knn", b, n, m, k, _ptr(unknown, _f32, "unknown"), _ptr(known, _f32, "known"),
              _ptr(dist2, _f32, "dist2"), _ptr(idx, _i32, "idx"), _stream())


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _lib.call("pn2_three_nn", b, n, m, _ptr(unknown, _f32, "unknown"), _ptr(known, _f32, "known"),
              _ptr(dist2, _f32, "dist2"), _ptr(idx, _i32, "idx"), _stream())


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _lib.call("pn2_three_interpolate", b, c, m, n, _ptr(points, _f32, "points"), _ptr(idx, _i32, "idx"),
              _ptr(weight, _f32, "weight"), _ptr(out, _f32, "out"), _stream())


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _lib.call("pn2_three_interpolate_grad", b, c, n, m, _ptr(grad_out, _f32, "grad_out"), _ptr(idx, _i32, "idx"),
              _ptr(weight, _f32, "weight"), _ptr(grad_points, _f32, "grad_points"), _stream())
