import sys, time, numpy as np
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, R+"/tests")
import pimd_tunneling_b200 as pk
from oracle_lib import Oracle, thermal_dimer_geometries, GOLDEN_GEOM_ANG
orc = Oracle()
pk.init()
# ---- ccpol energy + gradient
pes = pk.McmodMass("ccpol8sf").V_init()
orc.select("ccpol8sf")
x = thermal_dimer_geometries(64, seed=1)
t0=time.time(); v, g = pes.eval_batch(x); print("gpu eval", time.time()-t0)
vo, go, xo = orc.pes_eval(x)
print("golden kcal", pes.V(GOLDEN_GEOM_ANG.reshape(6,3).T/0.529177)*627.510)
print("V  max abs diff", np.abs(v-vo).max(), "bitwise equal:", np.array_equal(v, vo))
print("g  max rel diff", (np.abs(g-go).max(axis=(0,1))/np.abs(go).max(axis=(0,1))).max(), "bitwise equal:", np.array_equal(g, go))
xi = x.copy(order="F"); gi = pes.Vprime_batch_inplace(xi)
print("drift equal:", np.array_equal(xi, xo), "grad equal:", np.array_equal(gi, go))
# ---- 2d / 1d
for name in ("2dtest", "1d"):
    p2 = pk.McmodMass(name).V_init(); orc.select(name)
    xx = np.asfortranarray(np.random.default_rng(2).normal(0, 2.0, size=(p2.ndim, p2.natom, 1000)))
    v, g = p2.eval_batch(xx); vo, go, _ = orc.pes_eval(xx)
    print(name, "V equal", np.array_equal(v, vo), np.abs(v-vo).max(), "g equal", np.array_equal(g, go), np.abs(g-go).max())
