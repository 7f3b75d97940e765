"""time_ccpol.py [mode] [nbeads]: CUDA-event time of one strict CCpol gradient call over nbeads geometries (several passes on
the library's two streams), best and median of 7 -> one line.  PIMDK_LIB selects a variant library."""
import sys, os, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import torch
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check
from oracle_lib import thermal_dimer_geometries
pk.init(0)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8 * 32768
pes = pk.McmodMass("ccpol8sf").V_init()
check(lib().pimdk_set_mode(mode))
x = torch.from_numpy(np.ascontiguousarray(thermal_dimer_geometries(nb, seed=3).reshape(18, nb, order="F").T)).cuda()
g = torch.empty_like(x)
ts = []
for rep in range(9):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    check(lib().pimdk_pes_eval_dev(nb, 3, 6, x.data_ptr(), None, g.data_ptr()))
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts = sorted(ts[2:])
print("%s nbeads %d: best %.3f ms, median %.3f ms = %.4f ms per 32768-bead pass, %.3f M bead-gradients/s" % (
    os.path.basename(os.environ.get("PIMDK_LIB", "base")), nb, ts[0], ts[len(ts) // 2], ts[len(ts) // 2] / (nb / 32768.0), nb / ts[len(ts) // 2] / 1e3))
