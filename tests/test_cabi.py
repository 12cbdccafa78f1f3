"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares;
argument validation (no kernel launch involved) reports errors through the status code."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_ten_reference_ops():
    d = _declared()
    for op in ("furthest_point_sampling", "ball_query", "knn", "three_nn", "three_interpolate",
               "three_interpolate_grad", "group_points", "group_points_grad", "gather_points",
               "gather_points_grad"):
        assert "pn2_" + op in d


def test_library_exports_every_declared_symbol():
    from hotrack_b200 import _lib

    for name in _declared():
        assert hasattr(_lib.lib, name), "libpn2b200.so does not export %s" % name
    assert _lib.lib.pn2_version() >= 100


def test_every_bound_signature_is_declared():
    from hotrack_b200 import _lib

    d = set(_declared())
    for name in _lib.SIGNATURES:
        assert name in d, "%s is bound in _lib.py but not declared in include/" % name


def _prototypes():
    """name -> list of C parameter types, parsed from the headers."""
    protos = {}
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        for m in re.finditer(r"\bint\s+(pn2_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
            params = [q.strip() for q in m.group(2).replace("\n", " ").split(",")]
            protos[m.group(1)] = [] if params in ([""], ["void"]) else params
    return protos


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity and pointer / scalar kind of every ctypes binding against the C prototype it calls: a drift here is
    silent (ctypes passes whatever it is given) until a kernel reads a pointer where a size was meant."""
    from hotrack_b200 import _lib

    protos = _prototypes()
    for name, argtypes in _lib.SIGNATURES.items():
        params = protos[name]
        assert len(params) == len(argtypes), "%s: header has %d parameters, _lib.py binds %d" % (name, len(params), len(argtypes))
        for i, (c_decl, ct) in enumerate(zip(params, argtypes)):
            is_ptr = "*" in c_decl or "pn2_stream_t" in c_decl
            if ct is ctypes.c_void_p:
                assert is_ptr, "%s arg %d: bound as pointer, declared '%s'" % (name, i, c_decl)
            elif ct is ctypes.c_float:
                assert re.match(r"(const\s+)?float\s+\w+$", c_decl), "%s arg %d: bound as float, declared '%s'" % (name, i, c_decl)
            elif ct is ctypes.c_longlong:
                assert re.match(r"(const\s+)?long long\s+\w+$", c_decl), "%s arg %d: bound as long long, declared '%s'" % (name, i, c_decl)
            elif ct is ctypes.c_int:
                assert re.match(r"(const\s+)?int\s+\w+$", c_decl), "%s arg %d: bound as int, declared '%s'" % (name, i, c_decl)


def test_argument_errors_are_status_codes_not_exits():
    from hotrack_b200 import _lib

    st = _lib.lib.pn2_knn(1, 1, 1, 1 << 20, None, None, None, None, None)
    assert st != 0 and b"k >" in _lib.lib.pn2_last_error()
    st = _lib.lib.pn2_ball_query(-1, 1, 1, ctypes.c_float(0.1), 1, None, None, None, None)
    assert st != 0
    # empty problems are successful no-ops, as in the reference (sampling_gpu.cu:101)
    assert _lib.lib.pn2_furthest_point_sampling(0, 0, 0, None, None, None, None) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hotrack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "pn2_oracle" not in src and "libpn2_ref" not in src, f
