"""DFMA vs DMMA normal-mode transform on the C4 / C2 shapes: time, TFLOP/s, and agreement."""
import sys, os, time, ctypes, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check, hptr
pk.init(0)
L = lib()
for name, n, ndim, natom, nvec in (("C4", 512, 3, 6, 8192 * 18), ("C2", 256, 2, 1, 4096 * 2), ("C2/2", 256, 2, 1, 2048 * 2), ("C2x8", 256, 2, 1, 4096 * 2 * 8), ("n128", 128, 2, 1, 4096 * 2), ("n200", 200, 2, 1, 5000 * 2)):
    pes = pk.McmodMass("ccpol8sf" if natom == 6 else "2dtest").V_init()
    mass = np.ones(natom) * 1836.0
    check(L.pimdk_nm_setup(n, ndim, natom, hptr(mass), 12000.0 / (n + 1), 1.0))
    rng = np.random.default_rng(0)
    v = np.asfortranarray(rng.normal(size=(n, nvec)))
    out = {}
    for kind in (0, 3, 2):
        check(L.pimdk_set_gemm(kind))
        o = np.empty_like(v)
        check(L.pimdk_nm_transform(1, nvec, hptr(v), None, hptr(o)))
        out[kind] = o
        check(L.pimdk_profile(1)); check(L.pimdk_profile_reset())
        for _ in range(5): check(L.pimdk_nm_transform(1, nvec, hptr(v), None, hptr(o)))
        ms = ctypes.c_double(); c = ctypes.c_int64()
        check(L.pimdk_profile_get(b"gemm", ctypes.byref(ms), ctypes.byref(c)))
        check(L.pimdk_profile(0))
        t = ms.value / max(1, c.value)
        print("%s n=%d rows=%d  %s: %.3f ms per transform  %.2f TFLOP/s" % (name, n, nvec, {0: "DFMA", 3: "DMMA 128x128", 2: "DMMA 128x64"}[kind], t, 2.0 * n * n * nvec / (t * 1e-3) / 1e12))
    print("   bits equal to DFMA:", [bool(np.array_equal(out[0], out[k])) for k in (3, 2)])
check(L.pimdk_set_gemm(1))
