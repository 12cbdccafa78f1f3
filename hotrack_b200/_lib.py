"""ctypes binding of libpn2b200.so (C ABI: include/pn2b200.h).

There is no CPU or PyTorch fallback: if the CUDA library has not been built the
import fails loudly.  Build it with ``python hotrack_b200/build.py``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpn2b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "hotrack_b200: %s is missing -- the sm_100a kernels are not built. "
        "Run `python hotrack_b200/build.py` (needs nvcc); there is no fallback path." % LIB_PATH
    )

lib = ctypes.CDLL(LIB_PATH)

_i, _f, _p, _ll = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong

# name -> argtypes; every function returns int status (0 = ok)
SIGNATURES = {
    "pn2_furthest_point_sampling": [_i, _i, _i, _p, _p, _p, _p],
    "pn2_ball_query": [_i, _i, _i, _f, _i, _p, _p, _p, _p],
    "pn2_knn": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "pn2_three_nn": [_i, _i, _i, _p, _p, _p, _p, _p],
    "pn2_three_interpolate": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "pn2_three_interpolate_grad": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "pn2_group_points": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "pn2_group_points_grad": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "pn2_gather_points": [_i, _i, _i, _i, _p, _p, _p, _p],
    "pn2_gather_points_grad": [_i, _i, _i, _i, _p, _p, _p, _p],
    "pn2_adam_step": [ctypes.c_longlong, _p, _p, _p, _p, _f, _f, _f, _f, _f, _i, _p, _f, _p],
    # include/pn2b200_mlp.h
    "pn2_to_rows": [_i, _i, _i, _p, _p, _f, _p, _i, _p],
    "pn2_sa_build_rows": [_i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _p, _i, _i, _p, _p, _i, _p, _i, _p],
    "pn2_fp_build_rows": [_i, _i, _i, _p, _i, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _i, _p],
    "pn2_mlp_center": [_ll, _i, _i, _p, _i, _p, _p, _p, _p, _f, _i, _i, _p, _f, _i, _i, _p, _p, _p],
    "pn2_mlp_gemm_fwd": [_ll, _i, _i, _p, _i, _p, _p, _p, _p, _p, _i, _p, _p],
    "pn2_mlp_gemm_fwd_bn": [_ll, _i, _i, _p, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p,
                            _p, _p],
    "pn2_bn_finalize": [_i, _ll, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p],
    "pn2_bn_eval_affine": [_i, _p, _p, _p, _p, _p, _p, _f, _p, _p, _p],
    "pn2_pool_fwd": [_i, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p],
    "pn2_pool_bwd": [_i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p],
    "pn2_bn_bwd_coefs": [_i, _ll, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _p, _p, _p, _p, _p],
    "pn2_mlp_gemm_dgrad": [_ll, _i, _i, _p, _i, _p, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _p, _p],
    "pn2_mlp_gemm_wgrad": [_ll, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p, _p, _i, _p, _p, _p, _i, _p],
    "pn2_mlp_prep_weights": [_i, _i, _i, _p, _p, _p, _p],
    "pn2_mlp_prep_weights_multi": [_i, _p, _p],
    "pn2_sa_rows_bwd": [_i, _i, _i, _i, _p, _p, _i, _i, _p, _i, _i, _p, _i, _p],
    "pn2_fp_rows_bwd": [_i, _i, _i, _p, _p, _p, _i, _i, _p, _i, _i, _p, _p],
    # two-plane (hi + lo) forward rows
    "pn2_to_rows_x2": [_i, _i, _i, _p, _p, _f, _p, _p, _i, _p],
    "pn2_sa_build_rows_x2": [_i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _i, _p, _p, _i, _p],
    "pn2_fp_build_rows_x2": [_i, _i, _i, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p],
    "pn2_mlp_gemm_fwd_x2": [_ll, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p],
    "pn2_mlp_gemm_fwd_bn_x2": [_ll, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p,
                               _p, _p, _p, _p, _p, _p, _p],
    "pn2_pool_fwd_x2": [_i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p],
    "pn2_mlp_prep_weights_x2": [_i, _i, _i, _p, _p, _p, _p],
    # max-pool in the GEMM epilogue
    "pn2_mlp_gemm_fwd_pool": [_ll, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p],
    "pn2_mlp_gemm_fwd_bn_pool": [_ll, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p,
                                 _p, _p, _p, _p, _p, _p, _i, _p, _p, _p],
    "pn2_pool_finalize": [_i, _i, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p],
    # include/pn2b200_hand.h
    "pn2_kabsch_fwd": [_i, _i, _p, _i, _p, _p, _p, _p, _p],
    "pn2_kabsch_bwd": [_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p],
}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = _i
lib.pn2_version.restype = _i
lib.pn2_last_error.restype = ctypes.c_char_p
lib.pn2_launch_count.restype = ctypes.c_longlong


class Pn2Error(RuntimeError):
    pass


def check(status, name):
    if status != 0:
        raise Pn2Error("%s failed (status %d): %s" % (name, status, lib.pn2_last_error().decode()))


# Optional per-op device timing for bench.py's roofline line: when PROBE is a dict, every C-ABI
# call is bracketed by CUDA events on the current stream and (name, args, start, end) is recorded.
PROBE = None


# PN2_NVTX=1: an NVTX range around every C-ABI call (named after the entry point), for Nsight timelines.  Off by default:
# the step is ~250 launches and the ranges are host work.
NVTX = os.environ.get("PN2_NVTX", "") == "1"


def call(name, *args):
    if NVTX:
        import torch
        torch.cuda.nvtx.range_push(name)
        try:
            check(getattr(lib, name)(*args), name)
        finally:
            torch.cuda.nvtx.range_pop()
        return
    if PROBE is None:
        check(getattr(lib, name)(*args), name)
        return
    import torch

    # drain the stream first: with work still queued, the start event would fire while earlier kernels run and
    # the pair would time the queue, not this launch
    torch.cuda.current_stream().synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(getattr(lib, name)(*args), name)
    e.record()
    PROBE.setdefault(name, []).append((args, s, e))
