import sys, os, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+"/tests")
import torch
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check
from oracle_lib import thermal_dimer_geometries
pk.init(0)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 148*7*2
pes = pk.McmodMass("ccpol8sf").V_init()
check(lib().pimdk_set_mode(mode))
x = torch.from_numpy(np.ascontiguousarray(thermal_dimer_geometries(nb, seed=3).reshape(18, nb, order="F").T)).cuda()
g = torch.empty_like(x)
for rep in range(3):
    check(lib().pimdk_pes_eval_dev(nb, 3, 6, x.data_ptr(), None, g.data_ptr()))
torch.cuda.synchronize()
