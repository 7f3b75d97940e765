#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "malon or watmeth or water_methane" > $O/r2g_malon.log 2>&1; echo "malon rc=$?"; tail -15 $O/r2g_malon.log
timeout 300 python - <<'PY' 2>&1 | tail -12
import sys, time, numpy as np, torch
sys.path.insert(0, "tests")
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200._lib import lib, check
from oracle_lib import malon_geometries
pk.init(0)
pes = pk.McmodMass("malon").V_init()
nb = 148 * 128 * 8
x0 = malon_geometries(512, seed=1)
x = np.asfortranarray(np.tile(x0, (1, 1, nb // 512)))
xd = torch.from_numpy(np.ascontiguousarray(x.reshape(27, nb, order="F").T)).cuda()
g = torch.empty_like(xd); v = torch.empty(nb, dtype=torch.float64, device="cuda")
for what in ("grad", "energy"):
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        check(lib().pimdk_pes_eval_dev(nb, 3, 9, xd.data_ptr(), v.data_ptr() if what == "energy" else None, g.data_ptr() if what == "grad" else None))
        torch.cuda.synchronize(); dt = time.time() - t
    print("malon %s: %d geometries in %.2f ms = %.2f M/s" % (what, nb, dt * 1e3, nb / dt / 1e6))
PY
