"""ln(q/q0) of TI runs against the transfer-matrix value (tests/exact_pi.py) for several time steps and both thermostats"""
import os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import exact_pi
import pimd_tunneling_b200 as pk
from pimd_tunneling_b200.ti_driver import MCData, run_ti
pk.init(0)
a = np.array([[3.0], [0.0]]); b = np.array([[3.0 * np.cos(np.pi / 3)], [3.0 * np.sin(np.pi / 3)]])
n, beta = 15, 4.0
exact = exact_pi.log_ratio_2d(a[:, 0], b[:, 0], n, beta)
betan = beta / (n + 1)
for th in (2, 1):
    for dt, nmc in ((5e-3, 24000), (2e-3, 60000), (1e-3, 120000)):
        for seed in (1, 2):
            mc = MCData(n=n, beta=beta, NMC=nmc, imin=nmc // 10, dt=dt, nintegral=12, nrep=1024, thermostat=th, ndim=2, natom=1,
                        Noutput=int(1.5 / dt), seed=seed)
            t = time.time()
            res = run_ti("2dtest", mc, a, b, [1.0])
            got = -betan * res["deltaA"]; se = betan * np.sqrt(res["sigmaA"] / mc.nrep)
            print("th %d dt %.0e seed %d: %.5f +/- %.5f exact %.5f  z=%.2f  (%.1fs)" % (th, dt, seed, got, se, exact, (got - exact) / se, time.time() - t), flush=True)
