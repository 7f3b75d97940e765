"""ctypes binding of libpimdk.so (include/pimdk.h).  There is no CPU path: if the CUDA library is
missing or no sm_100a device is present, every compute call raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PIMDK_LIB", os.path.join(_HERE, "libpimdk.so"))
DATA_DIR = os.path.join(_HERE, "data")

_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_dbl = ctypes.c_double
_pd = ctypes.c_void_p  # double* (host numpy or device pointer)
_pi = ctypes.c_void_p

ERRORS = {1: "EINVAL", 2: "ENODEV", 3: "ECUDA", 4: "EDATA", 5: "ENAN", 6: "ENOCONV"}


class PimdkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pimdk error %s: %s" % (ERRORS.get(code, code), msg))
        self.code = code


_SIGS = {
    "pimdk_init": [_i64, ctypes.c_char_p],
    "pimdk_finalize": [],
    "pimdk_set_stream": [ctypes.c_void_p],
    "pimdk_set_mode": [_i64],
    "pimdk_set_fused": [_i64],
    "pimdk_set_gemm": [_i64],
    "pimdk_pes_select": [ctypes.c_char_p, _pd, _i64],
    "pimdk_pes_info": [ctypes.POINTER(_i64), ctypes.POINTER(_i64)],
    "pimdk_pes_set_v0": [_dbl],
    "pimdk_pes_eval": [_i64, _i64, _i64, _pd, _pd, _pd],
    "pimdk_pes_vprime_inplace": [_i64, _i64, _i64, _pd, _pd],
    "pimdk_pes_eval_dev": [_i64, _i64, _i64, _pd, _pd, _pd],
    "pimdk_um_forceenergy": [_i64, _i64, _i64, _pd, _pd, _pd, _pd, _dbl, _i64, _pd, _pd],
    "pimdk_um_forceenergy_batch": [_i64, _i64, _i64, _i64, _pd, _pd, _pd, _pd, _dbl, _i64, _pd, _pd],
    "pimdk_pes_hessian": [_i64, _i64, _i64, _pd, _pd],
    "pimdk_um_hessian": [_i64, _i64, _i64, _pd, _pd, _dbl, _i64, _pd],
    "pimdk_detj": [_i64, _i64, _i64, _pd, _pd, _dbl, _i64, _pd, _pd],
    "pimdk_readhess_displace": [_i64, _i64, _i64, _pd, _pd, _dbl, _dbl, _u64, _i64, _pd],
    "pimdk_nm_setup": [_i64, _i64, _i64, _pd, _dbl, _dbl],
    "pimdk_nm_get": [_pd, _pd, _pd],
    "pimdk_nm_transform": [_i64, _i64, _pd, _pd, _pd],
    "pimdk_init_path": [_i64, _i64, _pd, _pd, _pd, _pd, _u64, _pi, _pd, _pd],
    "pimdk_propagate": [_i64, _i64, _pd, _pd, _pd, _pd, _pd, _dbl, _dbl, _i64, _i64, _i64, _i64, _u64, _pi, _pd],
    "pimdk_propagate_dev": [_i64, _i64, _pd, _pd, _pd, _pd, _pd, _dbl, _dbl, _i64, _i64, _i64, _i64, _u64, _pi, _pd],
    "pimdk_set_restart": [_i64, _i64],
    "pimdk_set_andersen_carry": [_i64],
    "pimdk_set_dhdrlimit": [_dbl, _i64, _pd, _pd, _pd, _i64, _pd],
    "pimdk_set_propagate_chunk": [_i64],
    "pimdk_get_dhdr_sums": [_i64, _pd],
    "pimdk_ti_partial_sums": [_i64, _pd, _pi, _i64, _i64, _dbl, _pd],
    "pimdk_ti_finish": [_i64, _pd, _pd, _dbl, _pd, _pd, _pd, _pd, _pd],
    "pimdk_gauleg": [_dbl, _dbl, _i64, _pd, _pd],
    "pimdk_comm_unique_id": [ctypes.c_void_p],
    "pimdk_comm_init": [_i64, _i64, ctypes.c_void_p],
    "pimdk_comm_finalize": [],
    "pimdk_comm_info": [ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(_i64)],
    "pimdk_ti_allreduce": [_i64, _pd],
    "pimdk_ti_reduce_dev": [_i64, _pd, _pi, _i64, _i64, _dbl, _pd],
    "pimdk_profile": [_i64],
    "pimdk_profile_get": [ctypes.c_char_p, ctypes.POINTER(_dbl), ctypes.POINTER(_i64)],
    "pimdk_profile_reset": [],
    "pimdk_fp64_peak": [ctypes.POINTER(_dbl)],
    "pimdk_selftest_math": [_i64, _i64, _pd, _pd],
    "pimdk_selftest_division": [ctypes.POINTER(_i64)],
    "pimdk_selftest_fastmath": [ctypes.POINTER(_i64)],
}

_lib = None


def lib():
    """Load libpimdk.so (once).  Raises if it has not been built — the product has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s not found: build it with `python -m pimd_tunneling_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        L.pimdk_last_error.restype = ctypes.c_char_p
        L.pimdk_last_error.argtypes = []
        L.pimdk_last_nan_trajectory.restype = _i64
        L.pimdk_launch_count.restype = _i64
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PimdkError(rc, lib().pimdk_last_error().decode())


def hptr(arr):
    """host pointer of a C/F-contiguous float64 / int64 numpy array (or None)"""
    if arr is None:
        return None
    assert arr.flags["C_CONTIGUOUS"] or arr.flags["F_CONTIGUOUS"]
    return arr.ctypes.data_as(ctypes.c_void_p)


def f64(a, shape=None):
    a = np.asfortranarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


_initialised = False


def init(device=-1, data_dir=None):
    """pimdk_init; data_dir defaults to the packed tables shipped in the package."""
    global _initialised
    check(lib().pimdk_init(device, (data_dir or DATA_DIR).encode()))
    _initialised = True


def ensure_init():
    if not _initialised:
        init()


def finalize():
    global _initialised
    if _lib is not None:
        _lib.pimdk_finalize()
    _initialised = False
    # selections made before finalize are gone with the context
    from . import mcmod_mass, verletint

    mcmod_mass.McmodMass._owner = None
    verletint.VerletInt._owner = None
