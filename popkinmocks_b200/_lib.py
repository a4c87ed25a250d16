"""ctypes binding of libpkm_b200.so - the C ABI declared in include/pkm_b200.h.

This is the only place the Python host code touches native code.  There is no
CPU fallback: if the library has not been built, or no CUDA device is visible,
the compute entry points raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PKM_B200_LIB: development override (kernel-variant builds); the product is csrc/libpkm_b200.so
LIB_PATH = os.environ.get("PKM_B200_LIB") or os.path.join(_HERE, "csrc", "libpkm_b200.so")

PKM_OK = 0
PKM_ERR_INVALID = -1
PKM_ERR_CUDA = -2
PKM_ERR_OOM = -3
PKM_ERR_NO_ZERO_BIN = -4

LAYOUT_CUBE = 0
LAYOUT_PIXEL_MAJOR = 1


class PkmError(RuntimeError):
    """A C-ABI call failed; `.status` holds the pkm_status code."""

    def __init__(self, status, message):
        super().__init__(f"libpkm_b200: {message} (status {status})")
        self.status = status


class CubeDesc(C.Structure):
    _fields_ = [
        ("nt", C.c_int32), ("nz", C.c_int32), ("nv", C.c_int32), ("n_fft", C.c_int32),
        ("v0_idx", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32),
        ("reserved", C.c_int32), ("dv", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
    ]


_P = C.c_void_p
_I32, _I64, _U32, _U64, _F64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double

# name -> (restype, argtypes); mirrors include/pkm_b200.h one to one
SIGNATURES = {
    "pkm_last_error": (C.c_char_p, []),
    "pkm_abi_version": (_I32, []),
    "pkm_device_count": (_I32, []),
    "pkm_plan_create": (_I32, [C.POINTER(CubeDesc), _P, _P, _P, C.POINTER(_P)]),
    "pkm_plan_destroy": (_I32, [_P]),
    "pkm_plan_set_workspace_limit": (_I32, [_P, _I64]),
    "pkm_plan_launch_count": (_I64, [_P]),
    "pkm_plan_set_profiling": (_I32, [_P, _I32]),
    "pkm_plan_kernel_times": (_I32, [_P, _P, _P]),
    "pkm_ybar_density": (_I32, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P, _I32, _P]),
    "pkm_ybar_density_into_cube": (_I32, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P, _I32, _I32, _P]),
    "pkm_device_alloc": (_I32, [_I32, _I64, C.POINTER(_P)]),
    "pkm_device_free": (_I32, [_P]),
    "pkm_ipc_export": (_I32, [_P, _P]),
    "pkm_ipc_open": (_I32, [_P, _I32, C.POINTER(_P)]),
    "pkm_ipc_close": (_I32, [_P]),
    "pkm_host_register": (_I32, [_P, _I64]),
    "pkm_host_unregister": (_I32, [_P]),
    "pkm_ybar_gaussian": (_I32, [_P, _P, _P, C.POINTER(_I64), _P, C.POINTER(_I64), _P, _I32, _I32,
                                 _I32, _I32, _P, _I32, _P]),
    "pkm_axpy_cubes": (_I32, [_I32, C.POINTER(_P), _P, _I64, _P, _I32, _P]),
    "pkm_logsumexp_cubes": (_I32, [_I32, C.POINTER(_P), _P, _I64, _P, _I32, _P]),
    "pkm_transpose_to_cube": (_I32, [_P, _I64, _I64, _P, _I32, _P]),
    "pkm_irfft_batch": (_I32, [_P, _P, _I64, _F64, _P, _P]),
    "pkm_irfft_naive": (_I32, [_P, _P, _I64, _F64, _P, _P]),
    "pkm_reduce_logp": (_I32, [_P, _P, _I64, _P, _P, _U32, _I32, _P, _P]),
    "pkm_reduce_logp_multi": (_I32, [_P, _P, _I64, _P, _P, _I32, _P, _P, C.POINTER(_P), _P]),
    "pkm_reduce_logp_from": (_I32, [_P, _P, _U32, _I64, _P, _P, _U32, _I32, _P, _P]),
    "pkm_conditional": (_I32, [_P, _P, _U32, _P, _U32, _I64, _P, _I32, _I32, _P, _P]),
    "pkm_moments_logp": (_I32, [_P, _P, _P, _I64, _I64, _I64, _P, _I32, _P]),
    "pkm_moments": (_I32, [_P, _P, _I64, _I64, _P, _I32, _P]),
    "pkm_noise": (_I32, [_P, _I64, _F64, _I32, _U64, _U64, _F64, _P, _P, _I32, _P]),
    "pkm_max": (_I32, [_P, _I64, _P, _I32, _P]),
    "pkm_spline_tables": (_I32, [_I32, _I32, _P, _P, _P]),
    "pkm_fold_tables": (_I32, [_I32, _I32, _F64, _P, _P, _P, _P]),
    "pkm_histogram5d": (_I32, [C.POINTER(_P), _P, _I64, C.POINTER(_P), C.POINTER(_I32), _P, _I32, _P]),
    "pkm_pixelate_gaussian": (_I32, [_P, _P, _P, _P, _I32, _I32, _I64, _I32, _I32, _P, _I32, _P]),
    "pkm_synth_log_density": (_I32, [_P, _I32, _I32, _I64, _I64, _I64, _I32, _U64, _F64, _I32, _P]),
    "pkm_disk_maps": (_I32, [_P, _P, _I32, _I32, _F64, _F64, _F64, _P, _I32, _F64, _P, _P, _I32, _P, _P, _P, _P, _P,
                             _F64, _I32, _P]),
    "pkm_stream_maps": (_I32, [_P, _P, _I32, _I32, _F64, _F64, _F64, _P, _P, _I32, _F64, _F64, _F64, _F64, _F64,
                               _F64, _F64, _P, _P, _F64, _I32, _P]),
    "pkm_assemble_txz": (_I32, [_P, _P, _P, _I32, _F64, _P, _I32, _P, _P, _I32, _I64, _I32, _P, _P, _I32, _P]),
    "pkm_ssp_operand": (_I32, [_P, _I32, _P, _P, _I32, _I32, _I32, _P, _P, _I32, _P]),
    "pkm_fp64_fma_peak": (_I32, [_I32, _I32, C.POINTER(_F64)]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C popkinmocks_b200/csrc` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`.  popkinmocks_b200 has no "
            "CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pkm_abi_version() != 1:
        raise ImportError("libpkm_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status):
    if status != PKM_OK:
        msg = load().pkm_last_error().decode("utf-8", "replace")
        if status == PKM_ERR_NO_ZERO_BIN:
            # same exception type the reference raises at base.py:822-823
            raise IndexError(msg)
        raise PkmError(status, msg)


def device_count():
    return int(load().pkm_device_count())


def require_gpu():
    if device_count() <= 0:
        raise PkmError(PKM_ERR_CUDA, "no CUDA device visible; popkinmocks_b200 has no CPU fallback")


# ----------------------------------------------------------------------------
# argument marshalling
# ----------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def as_f64(x):
    """C-contiguous float64 NumPy view/copy of a host array."""
    return np.ascontiguousarray(x, dtype=np.float64)


def ptr(x):
    """Raw address of a NumPy array or torch tensor (host or CUDA)."""
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def stream_ptr(stream):
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


# ----------------------------------------------------------------------------
# host-only helpers (usable without a GPU)
# ----------------------------------------------------------------------------
def spline_tables(nfi, nfo):
    A_inv = np.empty((nfi, nfi))
    idx = np.empty(nfo, dtype=np.int32)
    w = np.empty((nfo, 4))
    check(load().pkm_spline_tables(nfi, nfo, ptr(A_inv), ptr(idx), ptr(w)))
    return A_inv, idx, w


def fold_tables(nv, v0_idx, dv):
    nfi, ne, no = nv // 2 + 1, nv // 2 + 1, (nv - 1) // 2
    CR = np.empty((nfi, ne))
    CI = np.empty((nfi, max(no, 1)))
    ip = np.empty(ne, dtype=np.int32)
    im = np.empty(ne, dtype=np.int32)
    check(load().pkm_fold_tables(nv, v0_idx, float(dv), ptr(CR), ptr(CI), ptr(ip), ptr(im)))
    return CR, CI[:, :no], ip, im
