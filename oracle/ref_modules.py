"""Loader for the REFERENCE's own Python layer of this path (oracle/_ref/pyref, staged unmodified
from /root/reference by `make -C oracle pyref`).  TEST INFRASTRUCTURE ONLY.

``load(cuda=True)``  -> (pointnet_utils, backbones) modules of the reference running on the
                        reference's own CUDA kernels (oracle/_ref/libpn2_ref.so through
                        oracle/refshim/pointnet2_cuda.py);
``load(cuda=False)`` -> the same files imported with CUDA hidden, which selects the reference's
                        pure-torch CPU fallback (pointnet_utils.py:7-10,26-32,126-137,156-167):
                        the "reference CPU path" timed by bench.py's cpu_baseline.

The modules are imported under their own top-level names (``pointnet_utils``, ``backbones``,
``pointnet_lib``, ``pointnet2_cuda``) because that is how the reference files import each other,
then removed from ``sys.modules`` so nothing else in the process resolves those names to them.
"""
import importlib
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PYREF = os.path.join(_HERE, "_ref", "pyref")
SHIM = os.path.join(_HERE, "refshim")
_NAMES = ("pointnet_utils", "backbones", "pointnet_lib", "pointnet_lib.pointnet2_utils", "pointnet2_cuda")
_cache = {}


def available(cuda=True):
    ok = os.path.exists(os.path.join(PYREF, "pointnet_utils.py"))
    if cuda:
        ok = ok and os.path.exists(os.path.join(_HERE, "_ref", "libpn2_ref.so"))
    return ok


def load(cuda=True):
    if cuda in _cache:
        return _cache[cuda]
    if not available(cuda):
        raise ImportError("oracle/_ref is not built: run `make -C oracle ref pyref` where /root/reference exists")
    saved_mods = {k: sys.modules.pop(k) for k in _NAMES if k in sys.modules}
    saved_path = list(sys.path)
    real_avail = torch.cuda.is_available
    try:
        sys.path[:0] = [PYREF, SHIM]
        if not cuda:
            torch.cuda.is_available = lambda: False  # read once, at import: pointnet_utils.py:7
        pu = importlib.import_module("pointnet_utils")
        bb = importlib.import_module("backbones")
        assert pu.CUDA == bool(cuda and real_avail()), "reference imported with the wrong CUDA switch"
    finally:
        torch.cuda.is_available = real_avail
        sys.path[:] = saved_path
        for k in _NAMES:
            sys.modules.pop(k, None)
        sys.modules.update(saved_mods)
    _cache[cuda] = (pu, bb)
    return pu, bb
