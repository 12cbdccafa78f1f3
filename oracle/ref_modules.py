"""Loader for the REFERENCE's own Python layer of this path (oracle/_ref/pyref, staged unmodified
from /root/reference by `make -C oracle pyref`).  TEST INFRASTRUCTURE ONLY.

``load(cuda=True)``  -> (pointnet_utils, backbones) modules of the reference running on the
                        reference's own CUDA kernels (oracle/_ref/libpn2_ref.so through
                        oracle/refshim/pointnet2_cuda.py);
``load(cuda=False)`` -> the same files imported with CUDA hidden, which selects the reference's
                        pure-torch CPU fallback (pointnet_utils.py:7-10,26-32,126-137,156-167):
                        the "reference CPU path" timed by bench.py's cpu_baseline.
``load(backend="ours")`` -> the same UNMODIFIED files, but ``import pointnet2_cuda`` (pointnet_lib/pointnet2_utils.py:7)
                        resolves to hotrack_b200/dropin/pointnet2_cuda.py, i.e. libpn2b200.so: the drop-in proof
                        (tests/test_dropin_gpu.py).
``load_full(backend)`` -> additionally ``hand_network`` (HandTrackNet with its transformer / blocks / hand_utils imports;
                        ``chumpy``, which only the MANO layer's pickle loader needs, is stubbed).

The modules are imported under their own top-level names (``pointnet_utils``, ``backbones``,
``pointnet_lib``, ``pointnet2_cuda``, ...) because that is how the reference files import each other,
then removed from ``sys.modules`` so nothing else in the process resolves those names to them.
"""
import importlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PYREF = os.path.join(_HERE, "_ref", "pyref")
SHIM = os.path.join(_HERE, "refshim")
DROPIN = os.path.join(os.path.dirname(_HERE), "hotrack_b200", "dropin")
_NAMES = ("pointnet_utils", "backbones", "pointnet_lib", "pointnet_lib.pointnet2_utils", "pointnet2_cuda",
          "hand_network", "transformer", "blocks", "hand_utils", "utils", "configs", "configs.config", "pose_utils",
          "pose_utils.rotations", "third_party", "third_party.mano", "third_party.mano.our_mano", "chumpy")
_cache = {}


def available(cuda=True, full=False):
    ok = os.path.exists(os.path.join(PYREF, "pointnet_utils.py"))
    if full:
        ok = ok and os.path.exists(os.path.join(PYREF, "hand_network.py"))
    if cuda:
        ok = ok and os.path.exists(os.path.join(_HERE, "_ref", "libpn2_ref.so"))
    return ok


def _import(cuda, backend, full):
    key = (bool(cuda), backend, bool(full))
    if key in _cache:
        return _cache[key]
    if not available(cuda and backend == "ref", full):
        raise ImportError("oracle/_ref is not built: run `make -C oracle ref pyref` where /root/reference exists")
    saved_mods = {k: sys.modules.pop(k) for k in _NAMES if k in sys.modules}
    saved_path = list(sys.path)
    real_avail = torch.cuda.is_available
    try:
        sys.path[:0] = [PYREF, SHIM if backend == "ref" else DROPIN]
        if not cuda:
            torch.cuda.is_available = lambda: False  # read once, at import: pointnet_utils.py:7
        if full:
            sys.modules["chumpy"] = types.SimpleNamespace(Ch=object)  # our_mano.py:10; only its pickle loader uses it
        pu = importlib.import_module("pointnet_utils")
        bb = importlib.import_module("backbones")
        assert pu.CUDA == bool(cuda and real_avail()), "reference imported with the wrong CUDA switch"
        hn = importlib.import_module("hand_network") if full else None
        if cuda:
            want = SHIM if backend == "ref" else DROPIN
            got = os.path.dirname(os.path.abspath(sys.modules["pointnet2_cuda"].__file__))
            assert got == want, "pointnet2_cuda resolved to %s, expected %s" % (got, want)
    finally:
        torch.cuda.is_available = real_avail
        sys.path[:] = saved_path
        for k in _NAMES:
            sys.modules.pop(k, None)
        sys.modules.update(saved_mods)
    _cache[key] = (pu, bb, hn)
    return _cache[key]


def load(cuda=True, backend="ref"):
    pu, bb, _ = _import(cuda, backend, False)
    return pu, bb


def load_full(backend="ref"):
    """(pointnet_utils, backbones, hand_network) of the reference on ``backend`` ('ref' | 'ours'), CUDA."""
    return _import(True, backend, True)


def handtracknet_cfg(device, handframe="camera"):
    """The cfg dict HandTrackNet reads (hand_network.py:51-54, backbones.py:80), with the reference's own
    pointnet2_camera_shallow1.yml."""
    import yaml

    with open(os.path.join(PYREF, "configs", "pointnet_config", "pointnet2_camera_shallow1.yml")) as f:
        cam = yaml.safe_load(f)
    return {"pointnet": {"camera": cam}, "device": device, "network": {"handframe": handframe, "backbone_out_dim": 384}}
