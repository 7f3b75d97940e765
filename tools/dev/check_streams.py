"""Multi-pass CCpol gradient call (passes of <= 32768 geometries, dealt alternately to two streams, PIMDK_CCPOL_STREAMS): every replica of a tiled batch must carry the bits of a separate single-pass call."""
import sys, os, numpy as np
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import torch
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check
from oracle_lib import thermal_dimer_geometries
pk.init(0)
pes = pk.McmodMass("ccpol8sf").V_init()
nd, rep = 1000, (int(sys.argv[1]) if len(sys.argv) > 1 else 70)
x0 = np.ascontiguousarray(thermal_dimer_geometries(nd, seed=5).reshape(18, nd, order="F").T)   # (nd, 18)
xs = torch.from_numpy(x0.copy()).cuda(); gs = torch.empty_like(xs)
check(lib().pimdk_pes_eval_dev(nd, 3, 6, xs.data_ptr(), None, gs.data_ptr()))
xb = torch.from_numpy(np.tile(x0, (rep, 1))).cuda(); gb = torch.empty_like(xb)
for it in range(2):   # twice: the second call reuses the streams and the staging halves
    gb.zero_()
    check(lib().pimdk_pes_eval_dev(nd * rep, 3, 6, xb.data_ptr(), None, gb.data_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(gb.view(rep, nd, 18), gs.unsqueeze(0).expand(rep, nd, 18)), "multi-pass gradient differs from the single-pass call"
assert torch.isfinite(gs).all()
pk.finalize(); print("check_streams ok:", nd * rep, "geometries in passes of at most 32768, bit-identical to the single-pass call")
