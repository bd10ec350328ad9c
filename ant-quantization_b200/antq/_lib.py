"""ctypes binding of libantq.so (C ABI: include/antq.h).

There is deliberately no fallback: if the shared library is missing and cannot be
built, importing this module raises, and every op raises on non-CUDA tensors.
"""
import ctypes
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
SO_PATH = os.path.join(_PKG, "csrc", "libantq%s.so" % os.environ.get("ANTQ_LIB_SUFFIX", ""))   # suffix: tuning builds only

F32, F16, BF16 = 0, 1, 2
FLAG_OVP, FLAG_FORCE_FLAT, FLAG_FORCE_ROWS, FLAG_FORCE_PU, FLAG_NO_PU, FLAG_FORCE_TILE = 1, 2, 4, 8, 16, 32
CB_WELLSEP, CB_STE_EXACT, CB_SYMMETRIC, CB_OVP_OK, CB_SYMX, CB_PU, CB_PU_UNIFORM, CB_PU_XC16, CB_PU_XCBF, CB_PU_E4M3 = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512
CB_PU_OVP = 1024
EINVAL, ENOTSUP, EALIGN = -1, -2, -3
MAX_GRID = 512
CODE_NONE = -1


class CodebookInfo(ctypes.Structure):
    _fields_ = [("n_entries", ctypes.c_int32), ("n_normal", ctypes.c_int32), ("n_levels", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("n_mag", ctypes.c_int32), ("mid", ctypes.c_int32),
                ("ovp_index", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("gmax", ctypes.c_float), ("vmax", ctypes.c_float), ("vmin", ctypes.c_float),
                ("lim", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _load():
    if not os.path.exists(SO_PATH):
        sys.path.insert(0, _PKG)
        try:
            import build as _build        # ant-quantization_b200/build.py (needs nvcc)
            _build.build()
        finally:
            sys.path.remove(_PKG)
    if not os.path.exists(SO_PATH):
        raise ImportError("libantq.so is missing and could not be built (%s)" % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    vp, i64, ci, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
    ip = ctypes.POINTER(CodebookInfo)
    L.antq_abi_version.restype = ci
    L.antq_build_info.restype = ctypes.c_char_p
    L.antq_error_string.restype = ctypes.c_char_p
    L.antq_error_string.argtypes = [ci]
    L.antq_codebook_bytes.restype = sz
    L.antq_codebook_prepare.argtypes = [vp, ci, vp, ci, vp, vp]
    L.antq_codebook_info_get.argtypes = [vp, ip, vp]
    L.antq_lut_nearest.argtypes = [vp, vp, vp, i64, ci, vp, vp]
    L.antq_fakequant.argtypes = [vp, vp, vp, vp, ci, i64, i64, ci, vp, ip, ci, vp]
    L.antq_fakequant_plan.argtypes = [ip, i64, i64, ci, ci, vp, vp, vp]
    L.antq_absmax.argtypes = [vp, vp, i64, i64, ci, vp]
    L.antq_fakequant_dynamic.argtypes = [vp, vp, vp, ctypes.c_float, i64, i64, ci, vp, ip, ci, vp]
    L.antq_mse_sweep.argtypes = [vp, vp, ci, vp, ci, vp, i64, i64, ci, vp, ci, vp]
    u32p = ctypes.POINTER(ctypes.c_uint32)
    L.antq_encode_p4.argtypes = [vp, vp, vp, ci, i64, i64, ci, vp, ip, ci, vp, vp]
    L.antq_decode_p4.argtypes = [vp, vp, vp, ci, i64, i64, ci, vp, ip, ci, vp]
    L.antq_backward_workspace_bytes.argtypes = [i64, i64, ci]
    L.antq_backward_workspace_bytes.restype = sz
    L.antq_fakequant_backward.argtypes = [vp, vp, vp, vp, ci, i64, i64, ci, ctypes.c_float, vp, vp, vp, sz, vp]
    L.antq_calibrate_workspace_bytes.argtypes = [i64, i64, ci, ci, ci]
    L.antq_calibrate_workspace_bytes.restype = sz
    L.antq_calibrate.argtypes = [vp, i64, i64, ci, ci, vp, vp, ci, ctypes.POINTER(vp), ctypes.POINTER(ip), ctypes.POINTER(ci),
                                 ci, vp, vp, vp, vp, sz, vp]
    L.antq_linear_p4.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, ci, vp, ip, ci, vp]
    L.antq_levels_e4m3.argtypes = [vp, vp, vp, i64, ci, vp, ip, vp]
    L.antq_linear_p4_fp8.argtypes = [vp, vp, vp, ip, vp, vp, vp, vp, i64, i64, i64, ci, vp, ip, ci, vp]
    L.antq_host_create.argtypes = [ctypes.POINTER(vp), ci, sz, ci]
    L.antq_host_destroy.argtypes = [vp]
    L.antq_host_destroy.restype = None
    L.antq_host_fakequant.argtypes = [vp, vp, vp, vp, ci, i64, i64, ci, vp, ci, vp, ci, ci]
    L.antq_host_fakequant_async.argtypes = [vp, vp, vp, vp, ci, i64, i64, ci, vp, ci, vp, ci, ci]
    L.antq_host_synchronize.argtypes = [vp]
    L.antq_host_last_launches.argtypes = [vp]
    for name in ("antq_codebook_prepare", "antq_codebook_info_get", "antq_lut_nearest", "antq_fakequant",
                 "antq_fakequant_plan", "antq_absmax", "antq_mse_sweep", "antq_host_create",
                 "antq_host_fakequant", "antq_host_fakequant_async", "antq_host_synchronize", "antq_host_last_launches",
                 "antq_encode_p4", "antq_decode_p4", "antq_fakequant_backward", "antq_calibrate", "antq_linear_p4", "antq_fakequant_dynamic", "antq_levels_e4m3", "antq_linear_p4_fp8"):
        getattr(L, name).restype = ci
    return L


lib = _load()
if lib.antq_abi_version() != 1:
    raise ImportError("libantq.so ABI version mismatch")


def check(status, what="antq"):
    if status != 0:
        raise RuntimeError("%s failed: %s (status %d)" % (what, lib.antq_error_string(status).decode(), status))
