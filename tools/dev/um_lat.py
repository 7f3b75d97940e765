import sys, os, time, ctypes, numpy as np
R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R)
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check, hptr
from bench import wells
pk.init(0)
pes = pk.McmodMass("2dtest").V_init()
a, b, mass = wells("2dtest")
n=1024
x0 = np.empty((n, 2, 1), order="F")
for i in range(n): x0[i] = a + (b - a) * i / (n - 1)
g = np.empty_like(x0); f = ctypes.c_double(0)
L = lib()
args = (n, 2, 1, hptr(x0), hptr(a), hptr(b), hptr(mass), 30.0/n, 1, ctypes.addressof(f), hptr(g))
for _ in range(50): check(L.pimdk_um_forceenergy(*args))
t0=time.perf_counter()
for _ in range(2000): L.pimdk_um_forceenergy(*args)
dt=time.perf_counter()-t0
print("raw C ABI call: %.1f us"%(dt/2000*1e6))
check(L.pimdk_profile(1)); check(L.pimdk_profile_reset())
for _ in range(200): L.pimdk_um_forceenergy(*args)
for fam in ("pes","um"):
    ms=ctypes.c_double(); c=ctypes.c_int64(); L.pimdk_profile_get(fam.encode(), ctypes.byref(ms), ctypes.byref(c)); print(fam, "%.1f us per call"%(ms.value/200*1e3), c.value)
im = pk.InstantonMod(pes, mass, 30.0, n, fixedends=True, rpi=True)
check(L.pimdk_profile(0))
t0=time.perf_counter()
for _ in range(2000): im.UMforceenergy(x0, a, b)
print("python wrapper call: %.1f us"%((time.perf_counter()-t0)/2000*1e6))
