import sys, os, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+"/tests")
import pimd_tunneling_b200 as pk
from bench import ti_path, wells
from pimd_tunneling_b200 import path as P
from oracle_lib import thermal_dimer_geometries
pk.init(0)
pes = pk.McmodMass("ccpol8sf").V_init()
x = thermal_dimer_geometries(37, seed=11)          # ragged: 37*36 energies, not a multiple of 32
v, g = pes.eval_batch(x)
h = pes.Vdoubleprime_batch(np.asfortranarray(x[..., :2].copy()))
for isurf in (1, 7, 8):   # Eckart embedding, SAPT-only (rigid/sweep stages skipped), potparts_old
    pk.McmodMass("ccpol8sf", isurf=isurf).V_init().eval_batch(x[..., :5])
pes = pk.McmodMass("ccpol8sf").V_init()
a, b, mass = wells("ccpol8sf")
for n, th in ((7, 2), (5, 1)):
    vi = pk.VerletInt(pes, n, mass, 160.0, dt=1e-3, NMC=2, Noutput=1, seed=7).init_nm()
    lam, path, spl = ti_path("ccpol8sf", a, b)
    xi = np.linspace(0.2, 0.8, 3)
    x0, p0 = vi.init_path(xi, lam, path, spl)
    bt, dbdl = P.endpoints(lam, path, spl, xi)
    (vi.propagate_pimd_pile if th == 2 else vi.propagate_pimd_nm)(x0, p0, a, bt, dbdl)
    im = pk.InstantonMod(pes, mass, 160.0, n)
    im.UMforceenergy(x0[..., 0], a, bt[..., 0]); im.detJ(x0[..., 0])
pes2 = pk.McmodMass("2dtest").V_init()
a2, b2, m2 = wells("2dtest")
for n in (33, 200):
    vi = pk.VerletInt(pes2, n, m2, 10.0, NMC=3, Noutput=2, seed=3).init_nm()
    lam, path, spl = ti_path("2dtest", a2, b2)
    xi = np.linspace(0.1, 0.9, 5)
    x0, p0 = vi.init_path(xi, lam, path, spl)
    bt, dbdl = P.endpoints(lam, path, spl, xi)
    vi.propagate_pimd_pile(x0, p0, a2, bt, dbdl); vi.propagate_pimd_nm(x0, p0, a2, bt, dbdl)
    if n == 33:
        vi.init_path(xi[:2], lam, path, spl, readhess=True)   # readhess thermal initialisation
        # the solid-angle loop's batched UMforceenergy (grid.y = polymer; 3 polymers, ragged against the 32-wide tiles)
        im2 = pk.InstantonMod(pes2, m2, 10.0, n)
        im2.UMforceenergy_batch(x0[..., :3], a2, bt[..., :3])
        pes2.Vdoubleprime(np.asfortranarray(a2))                # program crossover's transition-state Hessian
    # the chunked, copy-overlapped host-buffer path (chunks of 2, 2, 1 trajectories)
    from pimd_tunneling_b200._lib import check, lib
    check(lib().pimdk_set_propagate_chunk(2))
    vi.propagate_pimd_pile(x0, p0, a2, bt, dbdl); vi.propagate_pimd_nm(x0, p0, a2, bt, dbdl)
    check(lib().pimdk_set_propagate_chunk(0))
# 1D surface on the streamed Andersen path: the transforms' fused epilogues (model-surface gradient; kick + rotation + clocks)
pes1 = pk.McmodMass("1d").V_init()
a1, b1, m1 = wells("1d")
vi = pk.VerletInt(pes1, 130, m1, 10.0, NMC=3, Noutput=2, seed=4).init_nm()
lam, path, spl = ti_path("1d", a1, b1)
xi = np.linspace(0.1, 0.9, 3)
x0, p0 = vi.init_path(xi, lam, path, spl)
bt, dbdl = P.endpoints(lam, path, spl, xi)
vi.propagate_pimd_nm(x0, p0, a1, bt, dbdl); vi.propagate_pimd_pile(x0, p0, a1, bt, dbdl)
# the further plugin surfaces (row N4): ragged batches, Hessians, a short streamed propagation each
from oracle_lib import MALON_MASS, malon_geometries, watmeth_geometries
for name, geoms, masses in (("malon", malon_geometries(37, seed=3), list(MALON_MASS)), ("watmeth", watmeth_geometries(37, seed=3), [1837.0] * 17)):
    pz = pk.McmodMass(name).V_init()
    pz.eval_batch(geoms)
    pz.Vdoubleprime_batch(np.asfortranarray(geoms[..., :2].copy()))
    az, bz = np.asfortranarray(geoms[..., 0]), np.asfortranarray(geoms[..., 1])
    vz = pk.VerletInt(pz, 5, masses, 800.0, dt=1e-3, NMC=2, seed=2).init_nm()
    xz = np.asfortranarray(np.repeat(np.repeat(az[None, ...], 5, axis=0)[..., None], 3, axis=-1))
    pzz = np.zeros_like(xz, order="F")
    btz = np.asfortranarray(np.repeat(bz[..., None], 3, axis=-1))
    vz.propagate_pimd_pile(xz, pzz, az, btz, np.zeros_like(btz, order="F"))
ps = pk.McmodMass("so2").V_init()
ps.eval_batch(np.asfortranarray(np.random.default_rng(0).normal(size=(2, 1, 37)) + 14.0))
pk.finalize(); print("sanitize workload done")
