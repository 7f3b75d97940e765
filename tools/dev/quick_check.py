import sys, time, os, ctypes, numpy as np
R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, R+"/tests")
import torch
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check
from oracle_lib import Oracle, thermal_dimer_geometries
orc = Oracle().select("ccpol8sf")
pk.init()
pes = pk.McmodMass("ccpol8sf").V_init()
x = thermal_dimer_geometries(64, seed=1)
v, g = pes.eval_batch(x); vo, go, xo = orc.pes_eval(x)
print("bitwise: V", np.array_equal(v, vo), "grad", np.array_equal(g, go), "maxrel g", (np.abs(g-go).max()/np.abs(go).max()))
for mode in (0, 1):
    check(lib().pimdk_set_mode(mode))
    nb = 148*7*16
    xt = torch.from_numpy(np.ascontiguousarray(thermal_dimer_geometries(nb, seed=3).reshape(18, nb, order="F").T)).cuda()
    gt = torch.empty_like(xt)
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.time()
        check(lib().pimdk_pes_eval_dev(nb, 3, 6, xt.data_ptr(), None, gt.data_ptr()))
        torch.cuda.synchronize(); best = min(best, time.time() - t0)
    print("mode %d  nb=%d  grad time %.4f s  -> %.4e bead-grad/s" % (mode, nb, best, nb/best))
