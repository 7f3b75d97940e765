import sys, os, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+"/tests")
import pimd_tunneling_b200 as pk
from oracle_lib import thermal_dimer_geometries
pk.init(0)
pes = pk.McmodMass("ccpol8sf").V_init()
x = thermal_dimer_geometries(3, seed=11)
v, g = pes.eval_batch(x)
pk.finalize(); print("done")
