"""CPU: the C-ABI library loads and exports every symbol include/b200gan.h declares."""
import ctypes
import os
import re

from b200gan import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200gan.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s


def test_binding_table_matches_header():
    assert set(_capi.SIGNATURES) | {"b200_last_error"} == set(declared_symbols())


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    L = _capi.lib()
    assert L.b200_device_check() != 0
    assert b"CUDA" in L.b200_last_error() or b"device" in L.b200_last_error()


def test_route_is_host_only_logic():
    g = _capi.ConvGeom(N=8, H=32, W=32, Cin=3, Ho=16, Wo=16, Cout=200, k=5, stride=2, pad_t=1, pad_l=1)
    assert _capi.route(g, 0) == 2
    g.Cin, g.Cout = 200, 400
    assert _capi.route(g, 0) == 1 and _capi.route(g, 2) == 1
    g.Cin = 100
    try:
        _capi.route(g, 0)
        assert False
    except _capi.B200Error:
        pass
