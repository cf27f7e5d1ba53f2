"""ctypes binding of include/b200gan.h (the C-ABI drop-in boundary).

There is deliberately no fallback: if `lib/libb200gan.so` is missing or a call fails, we raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200GAN_LIB") or os.path.join(_HERE, "lib", "libb200gan.so")

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = range(5)
OPT_ADAM, OPT_RMSPROP, OPT_SGD, OPT_MOMENTUM, OPT_ADAGRAD, OPT_ADADELTA, OPT_FTRL, OPT_CENTERED_RMSPROP = range(8)


class ConvGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("N", "H", "W", "Cin", "Ho", "Wo", "Cout", "k", "stride", "pad_t", "pad_l")]


class Epilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("act", C.c_int), ("leak", C.c_float), ("mask_src", C.c_void_p),
                ("mask_kind", C.c_int), ("out_f32", C.c_int), ("accumulate", C.c_int),
                ("mask_bits", C.c_void_p), ("bits_out", C.c_void_p), ("bits_pitch", C.c_int)]


class TransposeEntry(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("tile_begin", C.c_longlong), ("T", C.c_int), ("A", C.c_int),
                ("B", C.c_int), ("reserved", C.c_int)]


class B200Error(RuntimeError):
    pass


_P, _I, _F, _LL, _ULL, _U = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_ulonglong, C.c_uint
_GP, _EP = C.POINTER(ConvGeom), C.POINTER(Epilogue)

# name -> argtypes; mirrors include/b200gan.h one to one (tests/test_abi.py checks the export list)
SIGNATURES = {
    "b200_set_tuning": [C.c_char_p, _I],
    "b200_conv2d_fprop": [_P, _P, _P, _P, _GP, _EP, _P, _LL, _P],
    "b200_conv2d_dgrad": [_P, _P, _P, _GP, _EP, _P, _LL, _P],
    "b200_conv2d_wgrad": [_P, _P, _P, _GP, _F, _P, _LL, _I, _P],
    "b200_conv2d_wgrad_bias": [_P, _P, _P, _P, _GP, _F, _P, _LL, _I, _P],
    "b200_conv2d_wgrad_folds_bias": [_GP, _I],
    "b200_conv2d_workspace_bytes": [_GP, _I],
    "b200_conv2d_route": [_GP, _I],
    "b200_conv2d_epilogue_bits": [_GP, _I, _I],
    "b200_gemv_rows": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "b200_outer_mask": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "b200_bn_sums": [_P, _P, _LL, _I, _P],
    "b200_bn_apply": [_P, _P, _P, _P, _LL, _I, _F, _I, _F, _P],
    "b200_bn_update_moving": [_P, _LL, _I, _P, _P, _F, _I, _P],
    "b200_bn_bwd": [_P, _P, _P, _P, _P, _LL, _I, _F, _P],
    "b200_maskmul": [_P, _P, _P, _LL, _I, _F, _P],
    "b200_affine_act": [_P, _I, _P, _I, _LL, _F, _F, _I, _F, _P],
    "b200_axpby": [_P, _I, _F, _P, _P, _I, _F, _P, _I, _LL, _P],
    "b200_mul_add": [_P, _P, _P, _P, _LL, _P],
    "b200_fill_f32": [_P, _LL, _F, _P],
    "b200_interp": [_P, _P, _P, _P, _I, _I, _P],
    "b200_rowscale": [_P, _P, _F, _F, _P, _I, _I, _P],
    "b200_dropout": [_P, _P, _P, _LL, _F, _P],
    "b200_instnorm_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "b200_instnorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "b200_layout_convert": [_P, _I, _P, _I, _I, _I, _I, _F, _F, _P],
    "b200_summary_stats": [_P, _I, _LL, _P, _P, _I, _P],
    "b200_montage": [_P, _I, _P, _I, _I, _I, _I, _I, _F, _F, _P],
    "b200_slice_cols": [_P, _LL, _I, _P, _LL, _I, _LL, _I, _P, _I, _F, _P],
    "b200_transpose_to_bf16": [_P, _I, _P, _I, _I, _I, _P],
    "b200_colsum": [_P, _P, _P, _LL, _I, _F, _P],
    "b200_reduce_sum": [_P, _I, _LL, _P, _F, _I, _P],
    "b200_wgan_loss": [_P, _I, _I, _F, _P, _P],
    "b200_eltloss": [_P, _I, _P, _LL, _I, _F, _F, _F, _P, _P, _I, _I, _F, _P],
    "b200_philox": [_P, _I, _LL, _ULL, _P, _U, _I, _P],
    "b200_optim_step": [_P, _P, _P, _P, _P, _P, _LL, _I, _F, _F, _F, _F, _F, _F, _I, _P, _P],
    "b200_transpose_batch": [_P, _I, _LL, _P],
    "b200_nccl_load": [C.c_char_p],
    "b200_nccl_version": [],
    "b200_nccl_unique_id": [_P],
    "b200_nccl_init": [_P, _I, _I, C.POINTER(C.c_void_p)],
    "b200_nccl_allreduce_f32": [_P, _P, _LL, _P],
    "b200_nccl_broadcast_f32": [_P, _P, _LL, _I, _P],
    "b200_nccl_destroy": [_P],
    "b200_device_check": [],
    "b200_abi_version": [],
}

_lib = None


def lib():
    """Load the shared library (once).  Raises B200Error when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error("CUDA extension not built: %s missing (run __graft_entry__.build())" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        L.b200_conv2d_workspace_bytes.restype = C.c_longlong
        L.b200_last_error.restype = C.c_char_p
        L.b200_last_error.argtypes = []
        L.b200_nccl_last_error.restype = C.c_char_p
        L.b200_nccl_last_error.argtypes = []
        _lib = L
    return _lib


def call(name, *args):
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        err = L.b200_nccl_last_error() if name.startswith("b200_nccl_") else L.b200_last_error()
        raise B200Error("%s failed (%d): %s" % (name, rc, err.decode()))
    return rc


def nccl_library_path():
    """The libnccl torch itself uses (the nvidia-nccl wheel), so that one NCCL build serves the whole process."""
    import glob
    import importlib.util
    spec = importlib.util.find_spec("nvidia")
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        hits = sorted(glob.glob(os.path.join(base, "nccl", "lib", "libnccl.so*")))
        if hits:
            return hits[0]
    return None


def workspace_bytes(geom, op):
    return int(lib().b200_conv2d_workspace_bytes(C.byref(geom), op))


def set_tuning(key, value):
    """b200_set_tuning: run-time override of a launch-planner knob (include/b200gan.h)."""
    if lib().b200_set_tuning(key.encode(), int(value)) != 0:
        raise B200Error(lib().b200_last_error().decode())


def epilogue_bits(geom, op, has_workspace):
    """True when that conv call reads / writes the sign bitmaps of b200_epilogue (tensor-core epilogues)."""
    return lib().b200_conv2d_epilogue_bits(C.byref(geom), op, int(bool(has_workspace))) == 1


def wgrad_folds_bias(geom, has_workspace):
    """True when b200_conv2d_wgrad_bias also produces the bias gradient for this geometry (image-side layers)."""
    return lib().b200_conv2d_wgrad_folds_bias(C.byref(geom), int(bool(has_workspace))) == 1


def route(geom, op):
    """1 = tensor-core path, 2 = small-channel SIMT path; raises if unsupported."""
    L = lib()
    rc = L.b200_conv2d_route(C.byref(geom), op)
    if rc < 0:
        raise B200Error("unsupported conv geometry: %s" % L.b200_last_error().decode())
    return rc
