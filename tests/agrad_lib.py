"""TEST INFRASTRUCTURE: host build of the analytic-gradient mathematics (pimd_tunneling_b200/csrc/ccpol_grad.cuh is
__host__ __device__) wired per bead by tests/agrad_host.cpp, so that the formulas the CUDA kernels run can be checked on
the CPU against the oracle's dual-number gradient."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_agrad_host.so")
SRCS = [os.path.join(ROOT, "tests", "agrad_host.cpp"), os.path.join(ROOT, "pimd_tunneling_b200", "csrc", "ccpol_tables.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "pimd_tunneling_b200", "csrc", f) for f in ("ccpol_grad.cuh", "ccpol_tables.h")] + \
    [os.path.join(ROOT, "include", "pimdk_detmath.h")]
_P = ctypes.POINTER(ctypes.c_double)


def build():
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        tmp = SO + ".tmp.%d" % os.getpid()
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-o", tmp] + SRCS, check=True)
        os.replace(tmp, SO)
    return SO


class AgradHost:
    def __init__(self, isurf=3, iemon=1):
        self.L = ctypes.CDLL(build())
        self.L.agh_last_error.restype = ctypes.c_char_p
        data = os.path.join(ROOT, "pimd_tunneling_b200", "data")
        if self.L.agh_load(data.encode(), isurf, iemon) != 0:
            raise RuntimeError(self.L.agh_last_error().decode())
        self.iemon = iemon

    def sapt(self, a9, b9):
        a9, b9 = np.ascontiguousarray(a9, dtype=np.float64), np.ascontiguousarray(b9, dtype=np.float64)
        v, g = ctypes.c_double(), np.empty(18)
        self.L.agh_sapt(a9.ctypes.data_as(_P), b9.ctypes.data_as(_P), ctypes.byref(v), g.ctypes.data_as(_P))
        return v.value, g

    def energy_gradient(self, x18):
        x = np.ascontiguousarray(x18, dtype=np.float64)
        v, g = ctypes.c_double(), np.empty(18)
        rc = self.L.agh_energy_gradient(x.ctypes.data_as(_P), self.iemon, ctypes.byref(v), g.ctypes.data_as(_P))
        if rc:
            raise RuntimeError("no convergence in indN_iter")
        return v.value, g
