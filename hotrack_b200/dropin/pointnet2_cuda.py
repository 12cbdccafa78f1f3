"""``pointnet2_cuda`` -- put THIS directory on ``sys.path`` (ahead of site-packages) and the reference's
``network/models/pointnet_lib/pointnet2_utils.py:7`` (``import pointnet2_cuda as pointnet2``) binds to libpn2b200.so
instead of the extension built from ``pointnet_lib/src``.  Nothing else of the reference changes: its seven
``autograd.Function``s, its SA / FP modules, backbones and HandTrackNet run unmodified on the sm_100a kernels
(tests/test_dropin_gpu.py does exactly that, against the same files on the reference's own kernels).

The ten functions (reference src/pointnet2_api.cpp:10-24) live in ``hotrack_b200.pointnet2_cuda``; this file only gives
them the module name the reference imports.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:  # make the package importable when only this directory was put on the path
    sys.path.append(_ROOT)

from hotrack_b200.pointnet2_cuda import (  # noqa: E402,F401
    ball_query_wrapper, furthest_point_sampling_wrapper, gather_points_grad_wrapper, gather_points_wrapper,
    group_points_grad_wrapper, group_points_wrapper, knn_wrapper, three_interpolate_grad_wrapper,
    three_interpolate_wrapper, three_nn_wrapper)
