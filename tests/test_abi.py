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
    assert set(_capi.SIGNATURES) | {"b200_last_error", "b200_nccl_last_error"} == set(declared_symbols())


def test_nccl_entry_points_load_the_library_and_fail_loudly_without_a_communicator():
    """b200_nccl_* (SURVEY 8b): libnccl is resolved at run time; no communicator -> error text, never a silent no-op."""
    L = _capi.lib()
    path = _capi.nccl_library_path()
    assert path and os.path.exists(path)
    assert L.b200_nccl_load(path.encode()) == 0
    assert L.b200_nccl_version() >= 20000
    assert L.b200_nccl_allreduce_f32(None, None, 16, None) != 0
    assert b"communicator" in L.b200_nccl_last_error()


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


def test_launch_planner_is_host_only_logic_and_tunable():
    """conv_plan (csrc/capi.cu): the headline's layers fill the machine and never split; pix2pix's inner layers ask
    for the fp32 partial image of a split-K launch; b200_set_tuning overrides the choice and rejects unknown keys."""
    c2 = _capi.ConvGeom(N=512, H=16, W=16, Cin=208, Ho=8, Wo=8, Cout=416, k=5, stride=2, pad_t=1, pad_l=1)
    c3 = _capi.ConvGeom(N=512, H=8, W=8, Cin=416, Ho=4, Wo=4, Cout=832, k=5, stride=2, pad_t=1, pad_l=1)
    e5 = _capi.ConvGeom(N=16, H=16, W=16, Cin=512, Ho=8, Wo=8, Cout=512, k=4, stride=2, pad_t=1, pad_l=1)
    for g in (c2, c3):
        assert _capi.workspace_bytes(g, 0) == 0 and _capi.workspace_bytes(g, 1) == 0
    assert _capi.workspace_bytes(e5, 0) >= 16 * 8 * 8 * 512 * 4
    _capi.set_tuning("tap_splits", 1)
    try:
        assert _capi.workspace_bytes(e5, 0) == 0
    finally:
        _capi.set_tuning("tap_splits", -1)
    assert _capi.workspace_bytes(e5, 0) > 0
    try:
        _capi.set_tuning("no_such_knob", 1)
        assert False
    except _capi.B200Error:
        pass


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Compile a probe against include/b200gan.h with the host C compiler and compare struct sizes and field
    offsets with the ctypes mirrors in _capi.py (a silent layout mismatch would corrupt every call)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        import pytest
        pytest.skip("no C compiler")
    probe = tmp_path / "probe.c"
    fields_e = ["bias", "act", "leak", "mask_src", "mask_kind", "out_f32", "accumulate", "mask_bits", "bits_out",
                "bits_pitch"]
    fields_g = ["N", "H", "W", "Cin", "Ho", "Wo", "Cout", "k", "stride", "pad_t", "pad_l"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200gan.h"', 'int main(void) {',
             '  printf("%zu %zu\\n", sizeof(b200_conv_geom), sizeof(b200_epilogue));']
    lines += ['  printf("%%zu\\n", offsetof(b200_epilogue, %s));' % f for f in fields_e]
    lines += ['  printf("%%zu\\n", offsetof(b200_conv_geom, %s));' % f for f in fields_g]
    lines += ['  return 0;', '}']
    probe.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    sizes, offs = [int(v) for v in out[:2]], [int(v) for v in out[2:]]
    assert sizes == [ctypes.sizeof(_capi.ConvGeom), ctypes.sizeof(_capi.Epilogue)]
    want = [getattr(_capi.Epilogue, f).offset for f in fields_e] + [getattr(_capi.ConvGeom, f).offset for f in fields_g]
    assert offs == want


def test_epilogue_bits_query_is_host_only_logic():
    g = _capi.ConvGeom(N=8, H=16, W=16, Cin=208, Ho=8, Wo=8, Cout=400, k=5, stride=2, pad_t=1, pad_l=1)
    assert _capi.epilogue_bits(g, 0, False) and _capi.epilogue_bits(g, 1, False)     # tensor-core route
    g = _capi.ConvGeom(N=8, H=32, W=32, Cin=3, Ho=16, Wo=16, Cout=208, k=5, stride=2, pad_t=1, pad_l=1)
    assert _capi.epilogue_bits(g, 0, True) and not _capi.epilogue_bits(g, 0, False)  # image side: GEMM route only
    assert not _capi.epilogue_bits(g, 1, True)                                         # col2im epilogue: no bitmaps
    assert _capi.abi_version() == 4 if hasattr(_capi, "abi_version") else True
