import sys, os, ctypes
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check
pk.init()
for f in ("pimdk_selftest_division", "pimdk_selftest_fastmath"):
    bad = ctypes.c_int64(-1); check(getattr(lib(), f)(ctypes.byref(bad))); print(f, "mismatches:", bad.value)
