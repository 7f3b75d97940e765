"""GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI (libpimdk.so) and is
compared with the CPU oracle on the same seeded inputs.

Tolerances (north_star: deterministic paths to 1e-10 relative FP64):
  * PES energy / gradient, strict mode: BIT-EXACT (the kernels evaluate the reference's operation order
    with the shared math policy; the FD gradient, eps=1e-4, amplifies any last-bit difference by ~1e3-1e4,
    so anything weaker than bit-exact cannot deliver 1e-10 on it — see DESIGN.md §Parity)
  * PES fast mode (FMA contraction): energy 1e-10 relative, gradient 2e-8 of max|grad| (the measured
    sensitivity of the reference's own FD gradient to contraction)
  * normal-mode transform, NVE / thermostatted steps, init_path, UM*: 1e-10 relative (max-norm)
"""
import json
import os

import numpy as np
import pytest

import ctypes

from oracle_lib import GOLDEN_GEOM_ANG, GOLDEN_VAL, GOLDEN_VALM, Oracle, random_dimer_geometries, thermal_dimer_geometries

ctypes_P = ctypes.POINTER(ctypes.c_double)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
M_O, M_H = 15.9949146221 * 1822.888486, 1.0078250321 * 1822.888486
DIMER_MASS = [M_O, M_H, M_H, M_O, M_H, M_H]
RTOL = 1e-10


@pytest.fixture(scope="module")
def pk():
    import pimd_tunneling_b200 as pk

    pk.init(0)
    yield pk
    pk.finalize()


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# ---------------------------------------------------------------- PES plugin (mcmod_mass) ---------
def test_ccpol_known_answer(pk):
    pes = pk.McmodMass("ccpol8sf").V_init()
    e = pes.V((GOLDEN_GEOM_ANG / 0.529177).reshape(6, 3).T) * 627.510
    # main_CCpol-8sf.f:182-183: 25.15660 printed by a build without -r8; with -r8 (repo makefile) 25.15658
    assert abs(e - 25.15660) < 3e-5
    pes0 = pk.McmodMass("ccpol8sf", params=[0]).V_init()   # iemonomer = 0
    e0 = pes0.V((GOLDEN_GEOM_ANG / 0.529177).reshape(6, 3).T) * 627.510
    assert abs(e0 - (-1.34676)) < 5.1e-6


# main_CCpol-8sf.f:180-183 (test_parameters): the reference's only golden vectors, 10 surfaces x {interaction, + monomers}
@pytest.mark.parametrize("isurf", range(1, 11))
def test_ccpol_all_surfaces_golden_and_bit_exact(pk, orc, isurf):
    """All ten surfaces of init_ccpol through the C ABI (Eckart or Radau embedding, potparts or potparts_old, four
    SAPT data files, with and without the CCpol-8s correction): the interaction energy of the known-answer geometry
    within the printed precision of val(isurf); with monomers within 3e-5 of valm(isurf) (printed by a build without
    -r8, see test_oracle.py); energies and finite-difference gradients of 40 thermal geometries bit-exact against
    the oracle."""
    xg = (GOLDEN_GEOM_ANG / 0.529177).reshape(6, 3).T
    pes0 = pk.McmodMass("ccpol8sf", isurf=isurf, iemonomer=0).V_init()
    assert abs(pes0.V(xg) * 627.510 - GOLDEN_VAL[isurf - 1]) < 5.1e-6
    pes = pk.McmodMass("ccpol8sf", isurf=isurf).V_init()
    assert abs(pes.V(xg) * 627.510 - GOLDEN_VALM[isurf - 1]) < 3e-5
    orc.load_ccpol(isurf, 1)
    orc.L.orc_pes_select(b"ccpol8sf")
    orc.ndim, orc.natom = 3, 6
    x = thermal_dimer_geometries(40, seed=100 + isurf)
    v, g = pes.eval_batch(x)
    vo, go, _ = orc.pes_eval(x)
    assert np.array_equal(v, vo) and np.array_equal(g, go), (np.abs(v - vo).max(), np.abs(g - go).max())
    pk.McmodMass("ccpol8sf").V_init()      # back to the plugin's surface
    orc.load_ccpol(3, 1)


def test_ccpol_far_separated_dimers_underflow_exactly(pk, orc):
    """Monomers 150 ... 400 bohr apart, mixed with bound dimers inside one 32-energy group of the sweep: the
    exponentials e^{-beta R} reach the policy's flush-to-zero range (below -708) and must agree with the oracle bit
    for bit, like everything else."""
    pes = pk.McmodMass("ccpol8sf").V_init()
    orc.select("ccpol8sf")
    x = thermal_dimer_geometries(24, seed=5)
    x[2, 3:, :] += np.linspace(150.0, 400.0, 24)[None, :]     # monomer B (atoms 4-6) moved along z
    x[2, 3:, ::3] -= np.linspace(150.0, 400.0, 24)[None, ::3]  # every third one back to the bound geometry
    v, g = pes.eval_batch(x)
    vo, go, _ = orc.pes_eval(x)
    assert np.array_equal(v, vo) and np.array_equal(g, go), (np.abs(v - vo).max(), np.abs(g - go).max())


@pytest.mark.parametrize("isurf,seed", [(3, 11), (3, 12), (1, 13), (10, 14)])
def test_ccpol_random_orientations_and_separations_bit_exact(pk, orc, isurf, seed):
    """Dimers in random relative orientations, 4.2 ... 14 bohr apart, monomers distorted by 0.12 bohr per coordinate: the
    repulsive wall, the small-argument branch of the damping functions, every quadrant of the embedding's angles and
    induction iterations of different lengths inside one warp — energies and finite-difference gradients bit for bit
    (Radau surfaces 3 and 10, Eckart surface 1)."""
    pes = pk.McmodMass("ccpol8sf", isurf=isurf).V_init()
    orc.load_ccpol(isurf, 1)
    orc.L.orc_pes_select(b"ccpol8sf")
    orc.ndim, orc.natom = 3, 6
    x = random_dimer_geometries(150, seed=seed)
    v, g = pes.eval_batch(x)
    vo, go, _ = orc.pes_eval(x)
    assert np.isfinite(vo).all() and np.isfinite(go).all()
    assert np.array_equal(v, vo) and np.array_equal(g, go), (np.abs(v - vo).max(), np.abs(g - go).max())
    assert vo.max() - vo.min() > 1e-3        # the batch does span the wall and the long range (hartree)
    pk.McmodMass("ccpol8sf").V_init()      # back to the plugin's surface
    orc.load_ccpol(3, 1)


@pytest.mark.parametrize("nbatch", [1, 7, 252, 1000])
def test_ccpol_energy_gradient_bit_exact(pk, orc, nbatch):
    pes = pk.McmodMass("ccpol8sf").V_init()
    orc.select("ccpol8sf")
    x = thermal_dimer_geometries(nbatch, seed=nbatch)
    v, g = pes.eval_batch(x)
    vo, go, xo = orc.pes_eval(x)
    assert np.array_equal(v, vo)
    assert np.array_equal(g, go)
    xi = x.copy(order="F")
    gi = pes.Vprime_batch_inplace(xi)
    assert np.array_equal(gi, go) and np.array_equal(xi, xo)   # in-place FD drift reproduced
    assert np.abs(xi - x).max() > 0.0


def test_ccpol_frozen_vectors_and_v0(pk):
    G = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))
    for name, shape in (("ccpol8sf", (3, 6)), ("2dtest", (2, 1)), ("1d", (1, 1))):
        pes = pk.McmodMass(name).V_init()
        d = G[name]
        x = np.array([float.fromhex(h) for h in d["x"]]).reshape(shape + (d["nbatch"],), order="F")
        v, g = pes.eval_batch(x)
        assert [float(t).hex() for t in v] == d["v"]
        assert [float(t).hex() for t in g.reshape(-1, order="F")] == d["grad"]
    pes = pk.McmodMass("ccpol8sf").V_init()
    x = thermal_dimer_geometries(3, seed=1)
    v = pes.V_batch(x)
    pes.set_V0(v[0])
    assert np.array_equal(pes.V_batch(x), v - v[0])
    pes.set_V0(0.0)


def test_ccpol_multi_pass_gradient_bit_identical(pk, orc):
    """A gradient call of more than 32768 geometries runs as several passes of the seven-kernel pipeline, dealt
    alternately to two streams with separate staging slices (PIMDK_CCPOL_STREAMS): 70 replicas of 1000 thermal
    geometries (70 000 geometries, 3 passes; twice, so that streams and slices are reused) must carry, replica by
    replica, the bits of a single-pass call — which is bit-exact against the oracle on a sample."""
    pes = pk.McmodMass("ccpol8sf").V_init()
    nd, rep = 1000, 70
    x = thermal_dimer_geometries(nd, seed=5)                       # (3, 6, nd)
    v1, g1 = pes.eval_batch(x)
    vo, go, _ = orc.pes_eval(x[..., :16])
    assert np.array_equal(g1[..., :16], go) and np.array_equal(v1[:16], vo)
    xb = np.asfortranarray(np.tile(x, (1, 1, rep)))
    for _ in range(2):
        vb, gb = pes.eval_batch(xb)
        assert np.array_equal(gb.reshape(3, 6, rep, nd), np.broadcast_to(g1[:, :, None, :], (3, 6, rep, nd)))
        assert np.array_equal(vb.reshape(rep, nd), np.broadcast_to(v1[None, :], (rep, nd)))


def test_ccpol_fast_mode_within_contraction_noise(pk, orc):
    from pimd_tunneling_b200._lib import check, lib

    pes = pk.McmodMass("ccpol8sf").V_init()
    orc.select("ccpol8sf")
    x = thermal_dimer_geometries(200, seed=9)
    vo, go, _ = orc.pes_eval(x)
    check(lib().pimdk_set_mode(1))
    try:
        v, g = pes.eval_batch(x)
    finally:
        check(lib().pimdk_set_mode(0))
    assert np.abs(v - vo).max() <= 1e-10 * np.abs(vo).max()
    gscale = np.abs(go).max(axis=(0, 1))
    assert (np.abs(g - go).max(axis=(0, 1)) / gscale).max() < 2e-8


@pytest.mark.parametrize("name,scale", [("1d", 1.5), ("2dtest", 2.5)])
def test_model_surfaces_bit_exact(pk, orc, name, scale):
    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    x = np.asfortranarray(np.random.default_rng(3).normal(0, scale, size=(pes.ndim, pes.natom, 4097)))
    v, g = pes.eval_batch(x)
    vo, go, _ = orc.pes_eval(x)
    assert np.array_equal(v, vo) and np.array_equal(g, go)
    # scalar forms of the plugin interface
    assert pes.V(x[:, :, 5]) == vo[5]
    assert np.array_equal(pes.Vprime(x[:, :, 5]), go[:, :, 5])
    gg, ee = pes.potforce(x[:, :, 6])
    assert ee == vo[6] and np.array_equal(gg, go[:, :, 6])


def test_edge_cases(pk):
    pes = pk.McmodMass("2dtest").V_init()
    v, g = pes.eval_batch(np.zeros((2, 1, 0), order="F"))     # empty batch
    assert v.size == 0 and g.size == 0
    with pytest.raises(pk.PimdkError):                        # wrong shape for the selected PES
        from pimd_tunneling_b200._lib import check, hptr, lib
        x = np.zeros((3, 6, 1), order="F")
        check(lib().pimdk_pes_eval(1, 3, 6, hptr(x), None, hptr(x.copy(order="F"))))
    pes = pk.McmodMass("ccpol8sf").V_init()
    x = thermal_dimer_geometries(2, seed=1)
    x[:, 4, 1] = x[:, 5, 1]                                   # coincident hydrogens -> NaN trap
    with pytest.raises(pk.PimdkError) as ei:
        pes.eval_batch(x)
    assert ei.value.code in (5, 6)


def test_nan_trap_reports_the_trajectory_single_shot_and_chunked(pk):
    """The reference STOPs with "NaN in pot propagation" (verletmodule.f90:533-536, 577-580); the library returns
    PIMDK_ENAN and the index of the first offending trajectory in the caller's batch — also when the host-buffer call
    is pipelined in chunks (the index is the batch index, not the chunk-local one) — and still returns the others."""
    from pimd_tunneling_b200._lib import check, lib

    pes = pk.McmodMass("2dtest").V_init()
    a, b = _wells("2dtest")
    n, ntraj = 160, 7
    vi = pk.VerletInt(pes, n, [1.0], 10.0, NMC=3, Noutput=10 ** 9, seed=5).init_nm()
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, 0.05, [1.0])
    x[3, 1, 0, 4] = np.nan
    try:
        for chunk in (10 ** 6, 3):
            check(lib().pimdk_set_propagate_chunk(chunk))
            with pytest.raises(pk.PimdkError) as ei:
                vi.propagate_pimd_pile(x, p, a, bt, dbdl)
            assert ei.value.code == 5 and "NaN" in str(ei.value)
            assert int(lib().pimdk_last_nan_trajectory()) == 4
    finally:
        check(lib().pimdk_set_propagate_chunk(0))


def test_division_by_small_integers(pk):
    """the damping series divides by 1..10 with a 3-instruction correctly rounded sequence: 0 mismatches vs '/'"""
    import ctypes
    from pimd_tunneling_b200._lib import check, lib

    bad = ctypes.c_int64(-1)
    check(lib().pimdk_selftest_division(ctypes.byref(bad)))
    assert bad.value == 0


def test_math_policy_host_equals_device(pk):
    """include/pimdk_detmath.h: the device form (branch-free exp, __fma_rn ...) gives the bits of the host form (what
    the oracle is built from) on the hot path's ranges AND at the edges (overflow, underflow, NaN, signed zero)."""
    import ctypes
    import subprocess
    import tempfile

    from pimd_tunneling_b200._lib import check, hptr, lib

    src = os.path.join(tempfile.mkdtemp(), "dm.c")
    with open(src, "w") as f:
        f.write('#include "pimdk_detmath.h"\n'
                "void dm(int kind, long n, const double* x, double* y){ for(long i=0;i<n;++i){ double v=x[i], r=v; switch(kind){\n"
                "case 0: r=pimdk_exp(v);break; case 1: r=pimdk_log(v);break; case 2: r=pimdk_sin(v);break; case 3: r=pimdk_cos(v);break;\n"
                "case 4: r=pimdk_acos(v);break; case 5: r=pimdk_tanh(v);break; case 6: r=pimdk_pow(v,-1.5);break;\n"
                "case 7: r=pimdk_pow(v,-3.0);break; case 8: r=pimdk_pow(v,0.66666666666666666);break; case 9: r=pimdk_atan(v);break;} y[i]=r; } }\n")
    so = src[:-2] + ".so"
    subprocess.run(["gcc", "-O2", "-march=native", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    "-o", so, src, "-lm"], check=True)
    H = ctypes.CDLL(so)
    rng = np.random.default_rng(12)
    edges = np.array([0.0, -0.0, 1.0, -1.0, 700.0, -700.0, 707.999, -707.999, 708.0, -708.0, -708.0000000001, 708.5, 709.0,
                      709.0000000001, 709.7, 709.9, 720.0, -720.0, -745.0, -746.0, 1e300, -1e300, np.inf, -np.inf, np.nan,
                      5e-324, -5e-324, 1e-310])
    args = {0: np.concatenate([edges, rng.uniform(-750, 720, 200000), rng.uniform(-60, 30, 200000)]),
            1: np.exp(rng.uniform(-40, 40, 100000)), 2: rng.uniform(-50, 50, 100000), 3: rng.uniform(-50, 50, 100000),
            4: np.concatenate([[1.0, -1.0, 0.5, -0.5, 0.0], rng.uniform(-1, 1, 100000)]), 5: rng.uniform(-25, 25, 100000),
            6: np.exp(rng.uniform(-10, 10, 100000)), 7: np.exp(rng.uniform(-10, 10, 100000)), 8: np.exp(rng.uniform(-10, 10, 100000)),
            9: np.concatenate([[0.0, -0.0, 1.0, -1.0, 1e300, -1e300, np.inf, -np.inf, np.nan], rng.uniform(-5, 5, 100000),
                               rng.uniform(-1e6, 1e6, 20000)])}
    for kind, x in args.items():
        x = np.ascontiguousarray(x, dtype=np.float64)
        yd, yh = np.empty_like(x), np.empty_like(x)
        check(lib().pimdk_selftest_math(kind, x.size, hptr(x), hptr(yd)))
        H.dm(kind, ctypes.c_long(x.size), x.ctypes.data_as(ctypes.c_void_p), yh.ctypes.data_as(ctypes.c_void_p))
        same = (yd.view(np.uint64) == yh.view(np.uint64)) | (np.isnan(yd) & np.isnan(yh))
        assert same.all(), (kind, x[~same][:5], yd[~same][:5], yh[~same][:5])


def test_fast_div_sqrt_bit_identical(pk):
    """branch-free IEEE division / sqrt sequences used in the site-site sums == built-ins, bit for bit"""
    import ctypes
    from pimd_tunneling_b200._lib import check, lib

    bad = ctypes.c_int64(-1)
    check(lib().pimdk_selftest_fastmath(ctypes.byref(bad)))
    assert bad.value == 0


# ---------------------------------------------------------------- module verletint ----------------
def test_normal_mode_tables_and_transform(pk, orc):
    pes = pk.McmodMass("1d").V_init()
    orc.select("1d")
    for n in (5, 64, 200):
        betan = 10.0 / (n + 1)
        vi = pk.VerletInt(pes, n, [1.3], 10.0, tau=0.7).init_nm()
        orc.nm_setup(n, [1.3], betan, 0.7)
        a, b = np.array([[-1.0]]), np.array([[0.8]])
        orc.init_nm(a, b)
        T, lam, bm, bv = orc.get_nm()
        assert np.array_equal(vi.transmatrix, T) and np.array_equal(vi.lam, lam) and np.array_equal(vi.beadmass, bm)
        assert relmax(vi.beadvec(a, b), bv) < 1e-14
        v = np.asfortranarray(np.random.default_rng(n).normal(size=(n, 37)))
        q = vi.nmtransform_forward(v)
        assert relmax(q, T @ v) < RTOL
        assert relmax(vi.nmtransform_backward(q), v) < RTOL          # T*T = I
        bvn = np.asfortranarray(np.repeat(bv, 37, axis=1))
        qf = vi.nmtransform_forward(v, bvn)
        assert relmax(qf, T @ v - bvn) < RTOL
        assert relmax(vi.nmtransform_backward(qf, bvn), v) < RTOL


def test_transform_engines_give_identical_bits(pk):
    """FMA-pipe tile GEMM, DMMA with 128 x 64 and with 128 x 128 CTA tiles: same k-ordered fused multiply-add chain
    per output element, so the three engines must agree bit for bit (ragged sizes included)"""
    from pimd_tunneling_b200._lib import check, hptr, lib

    rng = np.random.default_rng(17)
    for n, nvec in ((64, 5), (129, 300), (512, 1000), (250, 131), (256, 8192), (200, 57)):
        pes = pk.McmodMass("1d").V_init()
        vi = pk.VerletInt(pes, n, [1.0], 10.0).init_nm()
        v = np.asfortranarray(rng.normal(size=(n, nvec)))
        bv = np.asfortranarray(rng.normal(size=(n, nvec)))
        outs = []
        for kind in (0, 1, 3):
            check(lib().pimdk_set_gemm(kind))
            outs.append((vi.nmtransform_forward(v, bv), vi.nmtransform_backward(v, bv)))
        check(lib().pimdk_set_gemm(1))
        for o in outs[1:]:
            assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])
        assert relmax(vi.nmtransform_backward(outs[0][0], bv), v) < RTOL


def _traj_inputs(pes, n, ntraj, a, b, sigma, mass, seed=5):
    from pimd_tunneling_b200 import path as P

    rng = np.random.default_rng(seed)
    nd, na = pes.ndim, pes.natom
    xi = 0.2 + 0.6 * np.arange(ntraj) / max(1, ntraj - 1)
    if pes.name == "ccpol8sf":
        lam, path, spl = P.build_path(P.acceptor_switch_path(a, b, 9))
    else:
        pts = np.empty((2, nd, na), order="F")
        pts[0], pts[1] = a, b
        lam, path, spl = P.build_path(pts)
    bt, dbdl = P.endpoints(lam, path, spl, xi)
    x = np.empty((n, nd, na, ntraj), order="F")
    for t in range(ntraj):
        for k in range(n):
            x[k, :, :, t] = a + (bt[..., t] - a) * (k + 1) / (n + 1) + rng.normal(0, sigma, size=(nd, na))
    p = np.asfortranarray(rng.normal(0, 1.0, size=x.shape) * np.sqrt(np.asarray(mass))[None, None, :, None] * 0.02)
    return x, p, bt, dbdl, (lam, path, spl, xi)


CASES = [
    # name, n, ntraj, steps, thermostat, beta, mass, sigma, Noutput, cayley, gamma
    ("1d", 16, 5, 50, 2, 10.0, [1.0], 0.05, 100000, False, 1.0),
    ("1d", 16, 5, 50, 1, 10.0, [1.0], 0.05, 7, False, 1.0),
    ("1d", 33, 3, 40, 2, 10.0, [1.0], 0.05, 100000, True, 1.0),
    ("2dtest", 24, 4, 50, 2, 10.0, [1.0], 0.05, 100000, False, 1.0),
    ("2dtest", 24, 4, 50, 1, 10.0, [1.0], 0.05, 5, False, 1.0),
    ("2dtest", 24, 4, 100, 2, 10.0, [1.0], 0.05, 100000, False, 0.0),     # NVE (gamma=0)
    ("2dtest", 24, 4, 100, 1, 10.0, [1.0], 0.05, 10 ** 9, False, 1.0),   # NVE (no Andersen collision)
    ("ccpol8sf", 8, 3, 10, 2, 200.0, DIMER_MASS, 0.01, 100000, False, 1.0),
    ("ccpol8sf", 8, 3, 10, 1, 200.0, DIMER_MASS, 0.01, 3, False, 1.0),
    ("ccpol8sf", 8, 2, 20, 2, 200.0, DIMER_MASS, 0.01, 100000, False, 0.0),  # NVE
]


def _wells(name):
    import sys
    sys.path.insert(0, ROOT)
    from bench import wells
    return wells(name)[:2]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-n%d-th%d-%s%s" % (c[0], c[1], c[4], "cayley" if c[9] else "exact", "-nve" if c[10] == 0 or c[8] > 10 ** 8 else ""))
def test_propagate_matches_oracle(pk, orc, case):
    name, n, ntraj, steps, thermostat, beta, mass, sigma, Noutput, cayley, gamma = case
    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    a, b = _wells(name)
    vi = pk.VerletInt(pes, n, mass, beta, dt=1e-3, gamma=gamma, NMC=steps, Noutput=Noutput, cayley=cayley, seed=4242).init_nm()
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, sigma, mass)
    gid = np.arange(ntraj, dtype=np.int64) * 3 + 11
    fn = vi.propagate_pimd_pile if thermostat == 2 else vi.propagate_pimd_nm
    xg, pg, dg = fn(x, p, a, bt, dbdl, traj_gid=gid)
    for t in range(ntraj):
        orc.nm_setup(n, mass, vi.betan, 1.0, gamma, 1e-3, cayley, True)
        orc.init_nm(a, bt[..., t])
        orc.set_rng(4242, int(gid[t]))
        xo, po, do = orc.propagate(thermostat, x[..., t], p[..., t], dbdl[..., t], steps, 0, Noutput)
        assert relmax(xg[..., t], xo) < RTOL
        assert relmax(pg[..., t], po) < RTOL
        assert abs(dg[t] - do) <= RTOL * abs(do)


@pytest.mark.parametrize("name,n,mass", [("1d", 64, [1.0]), ("1d", 37, [1.3]), ("2dtest", 48, [1.0]), ("2dtest", 128, [1.0])])
@pytest.mark.parametrize("thermostat", [1, 2])
def test_fused_small_system_kernel_equals_streamed_path(pk, name, n, mass, thermostat):
    """the persistent warp-per-ring-polymer kernel and the streamed GEMM path give the same bits"""
    from pimd_tunneling_b200._lib import check, lib

    pes = pk.McmodMass(name).V_init()
    a, b = _wells(name)
    vi = pk.VerletInt(pes, n, mass, 10.0, NMC=40, imin=5, Noutput=6, seed=5).init_nm()
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, 9, a, b, 0.05, mass)
    gid = np.arange(9, dtype=np.int64) + 100
    fn = vi.propagate_pimd_pile if thermostat == 2 else vi.propagate_pimd_nm
    out = []
    for fused in (1, 0):
        check(lib().pimdk_set_fused(fused))
        out.append(fn(x, p, a, bt, dbdl, traj_gid=gid))
    check(lib().pimdk_set_fused(1))
    for u, v in zip(out[0], out[1]):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("name,n,ntraj,mass", [("2dtest", 256, 70, [1.0]), ("2dtest", 130, 3, [1.0]), ("1d", 130, 67, [1.3]), ("so2", 64, 5, [1.0])])
def test_transform_epilogues_equal_the_separate_kernels(pk, name, n, ntraj, mass):
    """Andersen steps on the streamed path: with the tensor-core engine the back-transform forms the model surface's bead
    gradient and the forward transform applies kick, rotation and collision clocks in their epilogues; with the FMA-pipe
    engine (pimdk_set_gemm(0)) the same step runs as separate transform, PES and update kernels.  Same arithmetic per element,
    so positions, momenta and estimator sums agree bit for bit (ragged row counts against the 128-row tiles, collisions on)."""
    from pimd_tunneling_b200._lib import check, lib

    if name == "so2":
        pes = pk.McmodMass("so2", params=[3.0, 2.5]).V_init()
        a = np.asfortranarray(np.array([[2.5], [0.0]]))
        b = np.asfortranarray(np.array([[2.5 * np.cos(1.0)], [2.5 * np.sin(1.0)]]))
    else:
        pes = pk.McmodMass(name).V_init()
        a, b = _wells(name)
    vi = pk.VerletInt(pes, n, mass, 10.0, NMC=25, imin=3, Noutput=4, seed=77).init_nm()
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, 0.05, mass)
    gid = np.arange(ntraj, dtype=np.int64) * 5 + 1
    out = []
    check(lib().pimdk_set_fused(0))         # streamed path also where the small-system kernel would take over
    try:
        for kind in (1, 0):
            check(lib().pimdk_set_gemm(kind))
            out.append(vi.propagate_pimd_nm(x, p, a, bt, dbdl, traj_gid=gid))
    finally:
        check(lib().pimdk_set_gemm(1))
        check(lib().pimdk_set_fused(1))
    for u, v in zip(out[0], out[1]):
        assert np.isfinite(u).all() and np.array_equal(u, v)


def test_partition_invariance_and_imin(pk):
    """results depend on the global trajectory id only, not on how trajectories are batched/sharded"""
    pes = pk.McmodMass("2dtest").V_init()
    a, b = _wells("2dtest")
    n, ntraj = 32, 12
    vi = pk.VerletInt(pes, n, [1.0], 10.0, NMC=30, imin=10, Noutput=6, seed=99).init_nm()
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, 0.05, [1.0])
    gid = np.arange(ntraj, dtype=np.int64)
    for fn in (vi.propagate_pimd_pile, vi.propagate_pimd_nm):
        xa, pa, da = fn(x, p, a, bt, dbdl, traj_gid=gid)
        perm = np.random.default_rng(1).permutation(ntraj)
        xb, pb, db = fn(x[..., perm], p[..., perm], a, bt[..., perm], dbdl[..., perm], traj_gid=gid[perm])
        assert np.array_equal(xa[..., perm], xb) and np.array_equal(pa[..., perm], pb) and np.array_equal(da[perm], db)
        lo = fn(x[..., :5], p[..., :5], a, bt[..., :5], dbdl[..., :5], traj_gid=gid[:5])
        hi = fn(x[..., 5:], p[..., 5:], a, bt[..., 5:], dbdl[..., 5:], traj_gid=gid[5:])
        assert np.array_equal(np.concatenate([lo[2], hi[2]]), da)
    # imin: the estimator averages steps imin+1..NMC only (verletmodule.f90:397,413)
    vi2 = pk.VerletInt(pes, n, [1.0], 10.0, NMC=30, imin=0, Noutput=6, seed=99).init_nm()
    _, _, d0 = vi2.propagate_pimd_pile(x, p, a, bt, dbdl, traj_gid=gid)
    assert not np.array_equal(d0, vi.propagate_pimd_pile(x, p, a, bt, dbdl, traj_gid=gid)[2])


@pytest.mark.parametrize("name,n,mass,beta,sigma", [("1d", 32, [1.0], 10.0, 0.05), ("2dtest", 160, [1.0], 10.0, 0.05),
                                                   ("ccpol8sf", 8, DIMER_MASS, 200.0, 0.01)])
@pytest.mark.parametrize("thermostat", [1, 2])
def test_chunked_host_propagate_equals_single_shot(pk, name, n, mass, beta, sigma, thermostat):
    """pimdk_propagate (host buffers) pipelines large batches in chunks of whole trajectories with the copies
    overlapped; every bit of x, p, dHdr and of the running sums must equal the single-shot call (11 trajectories in
    chunks of 3: a ragged last chunk; fused and streamed paths; restart = 2 continuing the sums)."""
    from pimd_tunneling_b200._lib import check, lib

    pes = pk.McmodMass(name).V_init()
    a, b = _wells(name)
    ntraj, steps = 11, 12 if name != "ccpol8sf" else 4
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, sigma, mass)
    gid = np.arange(ntraj, dtype=np.int64) * 3 + 5
    vi = pk.VerletInt(pes, n, mass, beta, dt=1e-3, NMC=steps, imin=1, Noutput=5, seed=77).init_nm()
    fn = vi.propagate_pimd_pile if thermostat == 2 else vi.propagate_pimd_nm
    try:
        check(lib().pimdk_set_propagate_chunk(10 ** 6))      # one shot
        ref = fn(x, p, a, bt, dbdl, traj_gid=gid)
        sums_ref = vi.last_sums.copy()
        check(lib().pimdk_set_propagate_chunk(3))            # chunks 3, 3, 3, 2
        got = fn(x, p, a, bt, dbdl, traj_gid=gid)
        assert all(np.array_equal(u, v) for u, v in zip(ref, got)) and np.array_equal(sums_ref, vi.last_sums)
        # restart = 2: the incoming dHdr holds running sums and is continued, chunk by chunk
        vi.restart, vi.restartnmc = 2, steps
        check(lib().pimdk_set_propagate_chunk(10 ** 6))
        ref2 = fn(ref[0], ref[1], a, bt, dbdl, traj_gid=gid, dHdr0=sums_ref.copy())
        check(lib().pimdk_set_propagate_chunk(3))
        got2 = fn(ref[0], ref[1], a, bt, dbdl, traj_gid=gid, dHdr0=sums_ref.copy())
        assert all(np.array_equal(u, v) for u, v in zip(ref2, got2))
    finally:
        check(lib().pimdk_set_propagate_chunk(0))
        check(lib().pimdk_set_restart(0, 0))


@pytest.mark.parametrize("name,n,mass,beta,sigma", [("2dtest", 32, [1.0], 10.0, 0.05), ("2dtest", 160, [1.0], 10.0, 0.05),
                                                   ("ccpol8sf", 8, DIMER_MASS, 200.0, 0.01)])
@pytest.mark.parametrize("thermostat", [1, 2])
def test_restart_protocol(pk, tmp_path, name, n, mass, beta, sigma, thermostat):
    """write_restart / restart = 1, 2 (verletmodule.f90:162-185, 199-206, 246-247, 387-394, 412-413;
    pimd_par.f90:328-370): (i) the file round-trips every bit; (ii) a run segmented every Noutput steps with
    restart files equals the unsegmented run; (iii) 12 steps, then restart = 2 for 8 more, equals 20 steps
    (same Philox counters; the only difference is the extra p <-> P transform at each segment boundary)."""
    pes = pk.McmodMass(name).V_init()
    a, b = _wells(name)
    ntraj, steps = 3, 20 if name != "ccpol8sf" else 10
    first = steps * 3 // 5
    # PILE has no other use for Noutput, so it can serve as the restart cadence; the Andersen collision clock
    # restarts with every call (verletmodule.f90:202), so there the comparison is made without collisions
    nout = 8 if thermostat == 2 else 10 ** 9
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, sigma, mass)
    gid = np.arange(ntraj, dtype=np.int64) + 7

    def make(NMC, imin=0, Noutput=nout):
        return pk.VerletInt(pes, n, mass, beta, dt=1e-3, NMC=NMC, imin=imin, Noutput=Noutput, seed=31).init_nm()

    vi = make(steps, imin=2)
    fn = vi.propagate_pimd_pile if thermostat == 2 else vi.propagate_pimd_nm
    x_ref, p_ref, d_ref = fn(x, p, a, bt, dbdl, traj_gid=gid)
    # (i)
    f = str(tmp_path / vi.restart_filename(3, 12))
    assert f.endswith("restart_proc3_12.xyz")
    vi.write_restart(f, x_ref[..., 1], p_ref[..., 1], 1234, vi.last_sums[1])
    xr, pr, dr, nr = vi.read_restart(f)
    assert np.array_equal(xr, x_ref[..., 1]) and np.array_equal(pr, p_ref[..., 1]) and dr == vi.last_sums[1] and nr == 1234
    assert open(f).read().split()[0] == str(pes.natom)
    # (ii) restart = 1, segments of 8 steps (PILE) / one segment (Andersen)
    d1 = tmp_path / "seg"
    d1.mkdir()
    vs = make(steps, imin=2, Noutput=8 if thermostat == 2 else nout)
    vs.restart = 1
    xs, ps, ds = vs.propagate_restartable(thermostat, x, p, a, bt, dbdl, traj_gid=gid, iproc=0, directory=str(d1))
    assert relmax(xs, x_ref) < RTOL and relmax(ps, p_ref) < RTOL and np.abs(ds - d_ref).max() <= RTOL * np.abs(d_ref).max()
    _, _, _, done = vs.read_restart(str(d1 / vs.restart_filename(0, 1)))
    assert done == steps
    # (iii) `first` steps with restart = 1, then restart = 2 for the rest, against one run of `steps` (imin = 0)
    v0 = make(steps)
    fn0 = v0.propagate_pimd_pile if thermostat == 2 else v0.propagate_pimd_nm
    x0, p0, d0 = fn0(x, p, a, bt, dbdl, traj_gid=gid)
    d2 = tmp_path / "cont"
    d2.mkdir()
    va = make(first, Noutput=nout if thermostat == 1 else 10 ** 6)
    va.restart = 1
    va.propagate_restartable(thermostat, x, p, a, bt, dbdl, traj_gid=gid, iproc=2, directory=str(d2))
    vb = make(steps - first, Noutput=nout if thermostat == 1 else 10 ** 6)
    vb.restart = 2
    xb, pb, db = vb.propagate_restartable(thermostat, np.zeros_like(x), np.zeros_like(p), a, bt, dbdl, traj_gid=gid, iproc=2,
                                          directory=str(d2))
    assert relmax(xb, x0) < RTOL and relmax(pb, p0) < RTOL and np.abs(db - d0).max() <= RTOL * np.abs(d0).max()


def test_init_path_matches_oracle(pk, orc):
    for name, n, beta, mass in (("2dtest", 40, 10.0, [1.0]), ("ccpol8sf", 12, 300.0, DIMER_MASS)):
        pes = pk.McmodMass(name).V_init()
        orc.select(name)
        a, b = _wells(name)
        vi = pk.VerletInt(pes, n, mass, beta, seed=31337).init_nm()
        _, _, bt, _, (lam, path, spl, xi) = _traj_inputs(pes, n, 4, a, b, 0.0, mass)
        gid = np.array([0, 5, 6, 2 ** 31 + 7], dtype=np.int64)
        xg, pg = vi.init_path(xi, lam, path, spl, traj_gid=gid)
        for t in range(4):
            orc.nm_setup(n, mass, vi.betan)
            orc.init_nm(a, bt[..., t])
            orc.set_rng(31337, int(gid[t]))
            xo, po = orc.init_path(float(xi[t]), lam, path, spl)
            assert np.array_equal(xg[..., t], xo)          # spline evaluation: same operation order
            assert relmax(pg[..., t], po) < RTOL


# ---------------------------------------------------------------- module instantonmod -------------
@pytest.mark.parametrize("name,n,beta,mass,fixedends", [("2dtest", 1024, 30.0, [1.0], True), ("2dtest", 64, 30.0, [1.0], False),
                                                        ("1d", 128, 10.0, [1.0], True), ("ccpol8sf", 16, 300.0, DIMER_MASS, True)])
def test_um_forceenergy_matches_oracle(pk, orc, name, n, beta, mass, fixedends):
    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    a, b = _wells(name)
    im = pk.InstantonMod(pes, mass, beta, n, fixedends=fixedends, rpi=True)
    orc.nm_setup(n, mass, im.betan, 1.0, 1.0, 1e-3, False, fixedends)
    rng = np.random.default_rng(8)
    x = np.empty((n, pes.ndim, pes.natom), order="F")
    if name == "ccpol8sf":
        from pimd_tunneling_b200 import path as P
        lam, path, spl = P.build_path(P.acceptor_switch_path(a, b, 9))
        xs, _ = P.endpoints(lam, path, spl, np.linspace(0, 1, n))
        for k in range(n):
            x[k] = xs[..., k] + rng.normal(0, 0.01, size=(3, 6))
    else:
        for k in range(n):   # rpi_ser.f90:157-163 linear interpolation start
            x[k] = a + (b - a) * k / (n - 1) + rng.normal(0, 0.02, size=a.shape)
    g, f = im.UMforceenergy(x, a, b)
    go, fo = orc.UMforceenergy(x, a, b)
    assert f == fo                                    # scalar accumulated in the reference's order
    assert np.array_equal(g, go)
    assert im.UM(x, a, b) == orc.UM(x, a, b)
    assert np.array_equal(im.UMprime(x, a, b), orc.UMprime(x, a, b))


def test_instanton_lbfgs_iterates_match(pk, orc):
    """config C3 in miniature: scipy's L-BFGS-B (same algorithm as the vendored setulb: m=8, factr=1e6,
    pgtol=1e-5, maxls=40) driven by GPU f/g and by oracle f/g produce the same iterate sequence."""
    from scipy.optimize import fmin_l_bfgs_b

    name, n, beta = "2dtest", 256, 30.0
    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    a, b = _wells(name)
    im = pk.InstantonMod(pes, [1.0], beta, n, fixedends=True, rpi=True)
    orc.nm_setup(n, [1.0], im.betan)
    x0 = np.empty((n, 2, 1), order="F")
    for k in range(n):
        x0[k] = a + (b - a) * k / (n - 1)

    def run(fg):
        its = []
        def f(xf):
            x = np.asfortranarray(xf.reshape((n, 2, 1), order="F"))
            g, e = fg(x)
            return e, g.reshape(-1, order="F").copy()
        xf, fmin, info = fmin_l_bfgs_b(f, x0.reshape(-1, order="F"), m=8, factr=1e6, pgtol=1e-5, maxls=40, maxiter=60,
                                       callback=lambda xk: its.append(xk.copy()))
        return np.array(its), fmin

    ig, fgpu = run(lambda x: im.UMforceenergy(x, a, b))
    io, fcpu = run(lambda x: orc.UMforceenergy(x, a, b))
    assert ig.shape == io.shape and relmax(ig, io) < RTOL and abs(fgpu - fcpu) <= RTOL * abs(fcpu)


@pytest.mark.parametrize("name,n,beta,mass", [("2dtest", 48, 30.0, [1.0]), ("ccpol8sf", 6, 300.0, DIMER_MASS)])
def test_um_forceenergy_batch_is_bit_identical_per_polymer(pk, name, n, beta, mass):
    """pimdk_um_forceenergy_batch (the solid-angle loop's batch, rpi_par.f90:252-255): every polymer's UM and gradient
    equal the single-polymer call bit for bit, whatever its neighbours in the batch are."""
    pes = pk.McmodMass(name).V_init()
    a, b = _wells(name)
    im = pk.InstantonMod(pes, mass, beta, n, fixedends=True, rpi=True)
    rng = np.random.default_rng(8)
    npoly = 5
    X = np.empty((n, pes.ndim, pes.natom, npoly), order="F")
    B = np.empty((pes.ndim, pes.natom, npoly), order="F")
    for k in range(npoly):
        B[..., k] = b + 0.02 * rng.normal(size=b.shape)
        for i in range(n):
            X[i, :, :, k] = a + (B[..., k] - a) * i / (n - 1) + 0.01 * rng.normal(size=a.shape)
    G, F = im.UMforceenergy_batch(X, a, B)
    for k in range(npoly):
        g1, f1 = im.UMforceenergy(X[..., k], a, B[..., k])
        assert f1 == F[k] and np.array_equal(g1, G[..., k])


def test_instanton_batch_takes_the_iterates_of_separate_runs(pk):
    """instanton_batch runs one L-BFGS-B optimisation per end point side by side, their f/g requests coalesced into
    batched GPU calls: path, action and iteration count of every optimisation equal those of a run on its own."""
    from scipy.optimize import fmin_l_bfgs_b

    name, n, beta, mass = "2dtest", 64, 30.0, [1.0]
    pes = pk.McmodMass(name).V_init()
    a, b = _wells(name)
    pes.set_V0(pes.V(a))
    im = pk.InstantonMod(pes, mass, beta, n, fixedends=True, rpi=True)
    x0 = np.empty((n, 2, 1), order="F")
    for i in range(n):
        x0[i] = a + (b - a) * i / (n - 1)
    ends = np.array([b + d for d in ([[0.0], [0.0]], [[0.05], [-0.02]], [[-0.1], [0.04]], [[0.0], [0.15]])])
    xs, fs, infos = im.instanton_batch(x0, a, ends)
    for k in range(ends.shape[0]):
        def fg(v, k=k):
            g_, f_ = im.UMforceenergy(v.reshape(x0.shape, order="F"), a, ends[k])
            return f_, g_.reshape(-1, order="F")
        x1, f1, i1 = fmin_l_bfgs_b(fg, x0.reshape(-1, order="F"), m=8, factr=1e6, pgtol=pes.eps2, maxls=40, maxiter=15000)
        assert f1 == fs[k] and np.array_equal(x1.reshape(x0.shape, order="F"), xs[..., k])
        assert i1["nit"] == infos[k]["nit"] and i1["funcalls"] == infos[k]["funcalls"]
    pes.set_V0(0.0)


def test_angular_sweep_water_dimer_pipeline(pk):
    """The solid-angle loop of rpi_par.f90:209-300 end to end on the water dimer at a toy size (6 beads, 2 angles per
    axis = 8 rotated end points, 25 L-BFGS-B iterations each): rotate_atoms, batched instanton, detJ, I(beta)."""
    from pimd_tunneling_b200.instantonmod import rotate_atoms

    pes = pk.McmodMass("ccpol8sf").V_init()
    a, b = _wells("ccpol8sf")
    pes.set_V0(pes.V(a))
    n = 6
    im = pk.InstantonMod(pes, DIMER_MASS, 3000.0, n, fixedends=True, rpi=True)
    x0 = np.empty((n, 3, 6), order="F")
    for i in range(n):
        x0[i] = a
    r = im.angular_sweep(x0, a, a, 2, cutofftheta=0.02, cutoffphi=0.02, maxiter=25)
    assert r["Ibeta"].shape == (8,) and np.all(np.isfinite(r["Ibeta"])) and np.all((r["Ibeta"] >= 0) & (r["Ibeta"] <= 1))
    assert abs(r["weight"].sum() - 0.02 ** 3) < 1e-15
    # the first end point by hand: eta(1), phi(1), theta(1) about axes 1, 3, 1
    w = rotate_atoms(rotate_atoms(rotate_atoms(a, 1, r["eta"][0]), 3, r["phi"][0]), 1, r["theta"][0])
    g1, f1 = im.UMforceenergy(r["x"][..., 0], a, w)
    assert f1 == r["UM"][0]
    pes.set_V0(0.0)


# ---------------------------------------------------------------- stochastic agreement ------------
def test_ti_averages_agree_with_oracle_within_error_bars(pk, orc):
    """north_star: stochastic TI averages must agree with the reference within combined statistical error bars.
    1D double well, PILE: the GPU run and the CPU oracle use DIFFERENT seeds (independent samples); the per-lambda
    means of dH/dlambda must agree within 4 combined standard errors, and so must Delta A."""
    from pimd_tunneling_b200 import path as P
    from pimd_tunneling_b200.ti_driver import MCData, run_ti

    mc = MCData(n=16, beta=6.0, NMC=1500, imin=300, dt=5e-3, nintegral=4, nrep=48, thermostat=2, ndim=1, natom=1, seed=2024)
    a, b = np.array([[-1.0]]), np.array([[1.0]])
    res = run_ti("1d", mc, a, b, [1.0])
    nrep_o = 24
    orc.select("1d")
    betan = mc.beta / (mc.n + 1)
    lam, path, spl = P.build_path(np.stack([a, b], axis=0))
    xint, dbd = P.endpoints(lam, path, spl, res["xi"])
    I = np.empty((mc.nintegral, nrep_o))
    for k in range(mc.nintegral):
        for r in range(nrep_o):
            orc.nm_setup(mc.n, [1.0], betan, 1.0, 1.0, mc.dt)
            orc.init_nm(a, xint[..., k])
            orc.set_rng(777, 100000 + k * nrep_o + r)
            x0, p0 = orc.init_path(float(res["xi"][k]), lam, path, spl)
            _, _, d = orc.propagate(2, x0, p0, dbd[..., k], mc.NMC, mc.imin)
            I[k, r] = d / betan ** 2
    mo, so = I.mean(axis=1), I.std(axis=1, ddof=1) / np.sqrt(nrep_o)
    mg, sg = res["mean"], np.sqrt(np.maximum(res["var"], 0) / (mc.nrep - 1))
    z = np.abs(mg - mo) / np.sqrt(so ** 2 + sg ** 2)
    assert np.all(z < 4.0), (z, mg, mo)
    dA_o = np.sum(res["weights"] * mo)
    s_dA = np.sqrt(np.sum(res["weights"] ** 2 * (so ** 2 + sg ** 2)))
    assert abs(res["deltaA"] - dA_o) < 4.0 * s_dA
    # symmetric double well: the integrand is antisymmetric about lambda = 1/2, so Delta A ~ 0 within noise
    assert abs(res["deltaA"]) < 5.0 * np.sqrt(np.sum(res["weights"] ** 2 * sg ** 2)) + 1e-12


@pytest.mark.parametrize("name,n,beta,thermostat", [("1d", 16, 3.0, 2), ("1d", 64, 10.0, 2), ("2dtest", 15, 4.0, 1), ("2dtest", 15, 4.0, 2)])
def test_ti_ratio_matches_exact_path_integral(pk, name, n, beta, thermostat):
    """Known-answer test for the whole TI path (init_path, PILE / Andersen propagation, estimator, statistics),
    independent of the oracle: q/q0 = exp(-betan DeltaA) is a ratio of n-bead discretised density-matrix elements,
    which tests/exact_pi.py evaluates deterministically by transfer matrices.  The GPU run must reproduce ln(q/q0)
    within 4.5 standard errors of its own estimate (and the standard error must be small enough to mean something:
    the 2D value is -1.15, the 1D values -0.34 and -0.0014).  tools/ti_exact_scan.py repeats this for three time
    steps, both thermostats and 1024 repetitions: 12 runs, |z| <= 2.6, mean z = -0.1 (dt = 5e-3 shows the integrator's
    O(dt^2) bias at the 2-sigma level with PILE, so the test runs at 2e-3)."""
    import exact_pi
    from pimd_tunneling_b200.ti_driver import MCData, run_ti

    if name == "1d":
        a, b = np.array([[-1.0]]), np.array([[1.0]])
        exact = exact_pi.log_ratio_1d(-1.0, 1.0, n, beta)
        mc = MCData(n=n, beta=beta, NMC=60000, imin=5000, dt=2e-3, nintegral=12, nrep=256, thermostat=thermostat, ndim=1,
                    natom=1, seed=31337)
        tol_se = 0.02
    else:
        a = np.array([[3.0], [0.0]])
        b = np.array([[3.0 * np.cos(np.pi / 3)], [3.0 * np.sin(np.pi / 3)]])
        exact = exact_pi.log_ratio_2d(a[:, 0], b[:, 0], n, beta)
        mc = MCData(n=n, beta=beta, NMC=60000, imin=5000, dt=2e-3, nintegral=12, nrep=256, thermostat=thermostat, ndim=2,
                    natom=1, Noutput=750, seed=4711)
        tol_se = 0.03
    res = run_ti(name, mc, a, b, [1.0])
    betan = beta / (n + 1)
    got = -betan * res["deltaA"]
    se = betan * np.sqrt(res["sigmaA"] / mc.nrep)          # pimd_par.f90:420: sqrt(sigmaA/nrep)
    print("ln(q/q0): run %.5f +/- %.5f, exact %.5f" % (got, se, exact))
    assert se < tol_se, se
    assert abs(got - exact) < 4.5 * se, (got, exact, se)
    assert abs(np.log(res["q_over_q0"]) - got) < 1e-12


def test_ti_along_a_curved_path_file_gives_the_same_ratio(pk, tmp_path):
    """q/q0 depends on the end points only, not on the path b(xi) between them: a TI run along a curved path read from a
    path.xyz file (read_xyz_frames -> read_path: arc-length lampath, natural splines, xint/dbdxi on the curve) must
    reproduce the same exact ln(q/q0) = -1.147 as the straight line.  The frames lie on the arc of radius 3 between the
    two wells; the second run refines that guess to the instanton first (instapath, read_path :953-990)."""
    import exact_pi
    from pimd_tunneling_b200.ti_driver import MCData, run_ti

    n, beta = 15, 4.0
    a = np.array([[3.0], [0.0]])
    b = np.array([[3.0 * np.cos(np.pi / 3)], [3.0 * np.sin(np.pi / 3)]])
    f = tmp_path / "path.xyz"
    ang = np.linspace(0.0, np.pi / 3, 11)
    f.write_text("".join("1\npoint %d\nX %.12f %.12f\n" % (i, 3.0 * np.cos(t), 3.0 * np.sin(t)) for i, t in enumerate(ang)))
    exact = exact_pi.log_ratio_2d(a[:, 0], b[:, 0], n, beta)
    betan = beta / (n + 1)
    for instapath, seed in ((False, 11), (True, 12)):
        mc = MCData(n=n, beta=beta, NMC=60000, imin=5000, dt=2e-3, nintegral=12, nrep=256, thermostat=2, ndim=2, natom=1,
                    instapath=instapath, seed=seed)
        res = run_ti("2dtest", mc, a, b, [1.0], path_file=str(f))
        got = -betan * res["deltaA"]
        se = betan * np.sqrt(res["sigmaA"] / mc.nrep)
        print("curved path%s: ln(q/q0) %.5f +/- %.5f, exact %.5f" % (" (instanton)" if instapath else "", got, se, exact))
        assert se < 0.03 and abs(got - exact) < 4.5 * se, (instapath, got, exact, se)


# ---------------------------------------------------------------- second derivatives (row N2) -----
def test_vdoubleprime_bit_exact(pk, orc):
    """Vdoubleprime of the three plugins against the oracle's literal restatement, including the in-place drift"""
    rng = np.random.default_rng(21)
    for name, scale in (("1d", 1.5), ("2dtest", 2.5)):
        pes = pk.McmodMass(name).V_init()
        orc.select(name)
        x = np.asfortranarray(rng.uniform(-scale, scale, size=(pes.ndim, pes.natom, 50)))
        xg = x.copy(order="F")
        h = pes.Vdoubleprime_batch(xg)
        for t in range(50):
            ho, xo = orc.Vdoubleprime(x[..., t])
            assert np.array_equal(h[..., t], ho) and np.array_equal(xg[..., t], xo)
    # 2D: symmetric by construction of the reference's formula; 1D: close to the analytic second derivative
    pes = pk.McmodMass("1d").V_init()
    h1 = pes.Vdoubleprime(np.array([[0.7]]))
    assert abs(h1[0, 0, 0, 0] - (12 * 0.7 ** 2 - 4)) < 1e-6
    pes = pk.McmodMass("ccpol8sf").V_init()
    orc.select("ccpol8sf")
    x = thermal_dimer_geometries(3, seed=5)
    xg = x.copy(order="F")
    h = pes.Vdoubleprime_batch(xg)
    for t in range(3):
        ho, xo = orc.Vdoubleprime(x[..., t])
        assert np.array_equal(xg[..., t], xo)
        assert np.array_equal(h[..., t], ho)
    hm = h[..., 0].reshape(18, 18, order="F")
    assert np.abs(hm - hm.T).max() < 2e-3 * np.abs(hm).max()   # FD-of-FD noise level, eps 1e-5 / 1e-4


@pytest.mark.parametrize("name,n,beta,mass,singlewell", [("2dtest", 64, 30.0, [1.0], False), ("2dtest", 1024, 30.0, [1.0], False),
                                                         ("1d", 48, 10.0, [1.3], True), ("ccpol8sf", 6, 300.0, DIMER_MASS, False)])
def test_umhessian_and_detj(pk, orc, name, n, beta, mass, singlewell):
    """UMhessian band matrix bit-exact against the oracle; detJ eigenvalues against LAPACK's banded solver
    (scipy.linalg.eigvals_banded = DSBEVX/DSBEVD family, the routine the reference calls) on the oracle's matrix"""
    from scipy.linalg import eigvals_banded

    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    a, b = _wells(name)
    im = pk.InstantonMod(pes, mass, beta, n, fixedends=True, rpi=True)
    orc.nm_setup(n, mass, im.betan, 1.0, 1.0, 1e-3, False, True)
    rng = np.random.default_rng(3)
    x = np.empty((n, pes.ndim, pes.natom), order="F")
    if name == "ccpol8sf":
        from pimd_tunneling_b200 import path as P
        lam, path, spl = P.build_path(P.acceptor_switch_path(a, b, 9))
        xs, _ = P.endpoints(lam, path, spl, np.linspace(0, 1, n))
        for k in range(n):
            x[k] = xs[..., k]
    else:
        for k in range(n):
            x[k] = a + (b - a) * k / (n - 1) + rng.normal(0, 0.02, size=a.shape)
    band = im.UMhessian(x, singlewell)
    bo = orc.UMhessian(x, singlewell)
    assert np.array_equal(band, bo)
    eta = im.detJ(x, singlewell)
    ref = eigvals_banded(bo, lower=True)
    assert np.all(np.diff(eta) >= 0)
    assert np.abs(eta - ref).max() <= 1e-10 * np.abs(ref).max()
    if n <= 64:
        eta2, z = im.detJ(x, singlewell, eigvecs=True)
        assert np.abs(eta2 - ref).max() <= 1e-10 * np.abs(ref).max()
        N = n * pes.ndof
        dense = np.zeros((N, N))
        for c in range(N):
            for r in range(c, min(N, c + pes.ndof + 1)):
                dense[r, c] = dense[c, r] = bo[r - c, c]
        assert np.abs(dense @ z - z * eta2[None, :]).max() <= 1e-9 * np.abs(ref).max()
        assert np.abs(z.T @ z - np.eye(N)).max() < 1e-10


def test_rpi_splitting_closes_the_instanton_calculation(pk, orc):
    """rpi_ser.f90:221-236, 350-381 through the GPU path (instanton by L-BFGS-B on the GPU gradient, two detJ calls,
    kink action) against the same formulas on the oracle's UM / UMhessian with LAPACK eigenvalues."""
    from scipy.linalg import eigvals_banded
    from scipy.optimize import fmin_l_bfgs_b

    name, n, beta, mass = "2dtest", 128, 30.0, [1.0]
    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    a, b = _wells(name)
    v0 = pes.V(a)                 # V0 = V(well1) (rpi_ser.f90:95): the wells sit at zero
    pes.set_V0(v0)
    orc.set_V0(v0)
    im = pk.InstantonMod(pes, mass, beta, n, fixedends=True, rpi=True)
    orc.nm_setup(n, mass, im.betan, 1.0, 1.0, 1e-3, False, True)
    x0 = np.empty((n, 2, 1), order="F")
    for i in range(n):
        x0[i] = a + (b - a) * i / (n - 1)

    def fg(v):
        g_, f_ = im.UMforceenergy(v.reshape(x0.shape, order="F"), a, b)
        return f_, g_.reshape(-1, order="F")

    xs, _, info = fmin_l_bfgs_b(fg, x0.reshape(-1, order="F"), m=8, factr=1e6, pgtol=1e-5, maxls=40, maxiter=5000)
    xt = np.asfortranarray(xs.reshape(x0.shape, order="F"))
    r = im.rpi_splitting(xt, a, b)
    xh = np.empty_like(xt)
    xh[:] = a.reshape(1, 2, 1)
    e0 = eigvals_banded(orc.UMhessian(xh, True), lower=True)
    e1 = eigvals_banded(orc.UMhessian(xt, False), lower=True)[1:]
    l0, l1 = np.sum(np.log(e0[e0 > 0])), np.sum(np.log(e1[e1 > 0]))
    phi = np.exp(0.5 * (l1 - l0))
    sk = im.betan * orc.UM(xt, a, b)
    theta = im.betan * np.exp(-sk) * np.sqrt(sk / (2.0 * 3.14159265358979)) / phi
    assert abs(r["lndetj0"] - l0) <= 1e-10 * abs(l0) and abs(r["lndetj"] - l1) <= 1e-10 * abs(l1)
    assert abs(r["s_kink"] - sk) <= 1e-12 * abs(sk)
    assert abs(r["delta"] - 2.0 * theta / im.betan) <= 1e-8 * abs(2.0 * theta / im.betan)
    assert np.isfinite(r["delta"]) and r["delta"] > 0.0 and r["s_kink"] > 0.0
    pes.set_V0(0.0)
    orc.set_V0(0.0)


@pytest.mark.parametrize("name,n,mass,beta", [("1d", 24, [1.3], 8.0), ("2dtest", 16, [1.0], 6.0)])
def test_readhess_displacement_matches_oracle_composition(pk, orc, name, n, mass, beta):
    """init_path's `readhess` branch (verletmodule.f90:49-88) on the GPU against the same steps composed from the oracle:
    UMhessian band -> LAPACK eigenvectors, Philox normals (stream 4), the literal displacement formula.  Eigenvectors are
    defined up to sign, so the comparison is made mode by mode: the GPU displacement, mass-weighted and projected on the
    LAPACK eigenvectors, must have the amplitudes |tempx_m| / sqrt(eta2_m) (modes 2..totdof with eta2 >= 0; nothing along
    mode 1), and the eigenvalues must agree."""
    import ctypes
    from scipy.linalg import eig_banded
    from pimd_tunneling_b200 import path as P

    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    a, b = _wells(name)
    vi = pk.VerletInt(pes, n, mass, beta, seed=2718).init_nm()
    orc.nm_setup(n, mass, vi.betan, 1.0, 1.0, 1e-3, False, True)
    lam, path, spl = P.build_path(np.stack([a, b], axis=0))
    x0, _ = vi.init_path([0.7], lam, path, spl, traj_gid=[9])
    x0 = np.asfortranarray(x0[..., 0])
    band = orc.UMhessian(x0.copy(), False)
    xd = x0.copy(order="F")                          # x as UMhessian leaves it (1D/2D Vdoubleprime: no drift)
    w, Z = eig_banded(band, lower=True)
    xg, eta = vi.readhess_displace(x0, traj_gid=9)
    N = n * pes.ndof
    assert np.abs(eta - w).max() <= 1e-9 * np.abs(w).max()
    orc.L.orc_normal.restype = ctypes.c_double
    orc.L.orc_normal.argtypes = [ctypes.c_ulonglong, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_ulonglong]
    tempx = np.sqrt(1.0 / beta) * np.array([orc.L.orc_normal(2718, 4, 0, 9, m) for m in range(N)])
    # UMhessian numbers a degree of freedom as ndof*(bead-1) + dof; natom = 1 here, so idof2 is the same index
    dx = (xg - xd).reshape(n, pes.ndof, order="F").reshape(-1) * np.sqrt(mass[0])
    c = Z.T @ dx
    expect = np.where(w >= 0.0, np.abs(tempx) / np.sqrt(np.where(w > 0, w, 1.0)), 0.0)
    expect[0] = 0.0
    assert np.abs(np.abs(c) - expect).max() <= 1e-7 * max(expect.max(), 1e-300), (np.abs(c)[:4], expect[:4])
    # through the mirror of init_path: same polymer, same global id -> same bits
    x1, _ = vi.init_path([0.7], lam, path, spl, traj_gid=[9], readhess=True)
    assert np.array_equal(x1[..., 0], xg)


def test_rpi_splitting_matches_the_analytic_instanton(pk):
    """Known answer for the instanton rows (UM, UMprime, L-BFGS-B on the GPU gradient, UMhessian, detJ, the closing
    formulas of `program rpi`), independent of the oracle: for V = (x^2-1)^2 and mass m the kink action is
    S = (4/3) sqrt(2m) and the one-loop splitting is 2 w sqrt(6 S/pi) exp(-S), w = sqrt(8/m).  With m = 20, n = 512,
    beta = 40 the ring-polymer values converge to these to 3e-5 and 1.2e-4."""
    from scipy.optimize import fmin_l_bfgs_b

    m, n, beta = 20.0, 512, 40.0
    pes = pk.McmodMass("1d").V_init()
    pes.set_V0(0.0)
    a, b = np.array([[-1.0]]), np.array([[1.0]])
    im = pk.InstantonMod(pes, [m], beta, n, fixedends=True, rpi=True)
    x0 = np.empty((n, 1, 1), order="F")
    for i in range(n):
        x0[i] = a + (b - a) * i / (n - 1)

    def fg(v):
        g_, f_ = im.UMforceenergy(v.reshape(x0.shape, order="F"), a, b)
        return f_, g_.reshape(-1, order="F")

    xs, _, info = fmin_l_bfgs_b(fg, x0.reshape(-1, order="F"), m=8, factr=1e6, pgtol=1e-7, maxls=40, maxiter=20000)
    r = im.rpi_splitting(np.asfortranarray(xs.reshape(x0.shape, order="F")), a, b)
    S = 4.0 / 3.0 * np.sqrt(2.0 * m)
    delta = 2.0 * np.sqrt(8.0 / m) * np.sqrt(6.0 * S / np.pi) * np.exp(-S)
    assert r["skipped"] == 0 and r["skipped0"] == 0
    assert abs(r["s_kink"] - S) < 1e-4 * S, (r["s_kink"], S)
    assert abs(r["delta"] - delta) < 5e-4 * delta, (r["delta"], delta)


def test_rpi_front_end_reproduces_the_analytic_instanton(pk):
    """`program rpi` through the native front end (namelist RPIDATA, linear-interpolation guess, instanton, V0 shift,
    single-well Q_0, closing formulas): the quartic double well with m = 20 again, now from a namelist."""
    from pimd_tunneling_b200.rpi_driver import read_rpidata, run_rpi

    rd = read_rpidata("&RPIDATA\n n=512, beta=40.0d0, ndim=1, natom=1, readpath=.false., fixedends=.true.\n/\n")
    assert (rd.n, rd.beta, rd.ndim, rd.readpath, rd.npoints) == (512, 40.0, 1, False, 10)
    pes = pk.McmodMass("1d")
    eps2 = pes.eps2
    pk.McmodMass.eps2 = 1e-7          # pgtol (module variable eps2 of mcmod_mass) for a converged kink
    try:
        r = run_rpi("1d", rd, np.array([[-1.0]]), np.array([[1.0]]), [20.0])
    finally:
        pk.McmodMass.eps2 = eps2
    S = 4.0 / 3.0 * np.sqrt(40.0)
    delta = 2.0 * np.sqrt(8.0 / 20.0) * np.sqrt(6.0 * S / np.pi) * np.exp(-S)
    assert abs(r["s_kink"] - S) < 1e-4 * S and abs(r["delta"] - delta) < 5e-4 * delta, (r["s_kink"], S, r["delta"], delta)
    assert r["Vpath"].min() >= 0.0 and abs(r["lampath"][-1] - 1.0) < 1e-15


def test_crossover_temperature(pk):
    """`program crossover`: V = (x^2-1)^2 has V''(0) = -4, so beta_c = 2 pi / sqrt(4/m) = pi sqrt(m); the 2D surface's
    saddle between two wells (on the ring, at the half angle) has exactly one negative eigenvalue."""
    from pimd_tunneling_b200.rpi_driver import crossover

    pes = pk.McmodMass("1d").V_init()
    bc, eta = crossover(pes, [[0.0]], [7.0])
    assert abs(bc - np.pi * np.sqrt(7.0)) < 1e-6 * bc and abs(eta[0] + 4.0 / 7.0) < 1e-6   # Vdoubleprime is a finite difference (mcmod_1d.f90:37-57)
    pes2 = pk.McmodMass("2dtest").V_init()
    # radial position of the saddle on the 30-degree ray: maximum along the ring direction, minimum radially
    r = np.linspace(2.0, 4.0, 4001)
    ts = np.zeros((2, 1, r.size), order="F")
    ts[0, 0], ts[1, 0] = r * np.cos(np.pi / 6), r * np.sin(np.pi / 6)
    rs = r[np.argmin(pes2.V_batch(ts))]
    bc2, eta2 = crossover(pes2, [[rs * np.cos(np.pi / 6)], [rs * np.sin(np.pi / 6)]], [1.0])
    assert eta2[0] < 0.0 < eta2[1] and np.isfinite(bc2)


def test_full_size_c4_step_is_partition_invariant(pk):
    """BASELINE config C4 at its full size (512 beads x 8192 trajectories, CCpol-8sf, PILE): one step of the whole
    batch, then 12 sampled trajectories re-run alone and in a different order.  Results are keyed by the global
    trajectory id, so every bit must agree — a size-independent property that exercises the chunked PES pipeline
    (128 passes of 32 768 beads), the big-batch GEMM tiling and the RNG addressing at production scale."""
    import sys
    sys.path.insert(0, ROOT)
    from bench import CONFIGS, ti_path, wells
    from pimd_tunneling_b200 import path as P

    cfg = CONFIGS["c4"]
    n, ntraj = cfg["n"], cfg["nintegral"] * cfg["nrep"]
    a, b, mass = wells("ccpol8sf")
    pes = pk.McmodMass("ccpol8sf").V_init()
    vi = pk.VerletInt(pes, n, mass, cfg["beta"], dt=1e-3, NMC=1, seed=20261017).init_nm()
    xi, _ = vi.gauleg(0.0, 1.0, cfg["nintegral"])
    gid = np.arange(ntraj, dtype=np.int64)
    il = gid // cfg["nrep"]
    lam, path, spl = ti_path("ccpol8sf", a, b)
    xint, dbdxi = P.endpoints(lam, path, spl, xi)
    bt, dbdl = np.asfortranarray(xint[:, :, il]), np.asfortranarray(dbdxi[:, :, il])
    x0, p0 = vi.init_path(xi[il], lam, path, spl, traj_gid=gid)
    x1, p1, d1 = vi.propagate_pimd_pile(x0, p0, a, bt, dbdl, traj_gid=gid)
    assert np.isfinite(d1).all() and np.isfinite(x1).all()
    pick = np.array([0, 1, 511, 512, 4095, 4096, 8191, 7000, 3333, 2, 6144, 1023])
    xs, ps, ds = vi.propagate_pimd_pile(x0[..., pick], p0[..., pick], a, bt[..., pick], dbdl[..., pick], traj_gid=gid[pick])
    assert np.array_equal(xs, x1[..., pick]) and np.array_equal(ps, p1[..., pick]) and np.array_equal(ds, d1[pick])
    # and the thermostat did act: different repetitions of one lambda point have decorrelated
    assert not np.array_equal(p1[..., 0] - p0[..., 0], p1[..., 1] - p0[..., 1])


# ---------------------------------------------------------------- BASELINE sizes against the oracle -----------------
BASELINE_CASES = [
    # name, n, ntraj, steps, thermostat, beta, mass, sigma, Noutput   (the bead counts of BASELINE.json's configs)
    ("2dtest", 256, 2, 20, 1, 10.0, [1.0], 0.05, 7),            # C2: Andersen, streamed path (tensor-core transform,
                                                                 #     paired-normal update, last-bead estimator)
    ("2dtest", 256, 2, 20, 2, 10.0, [1.0], 0.05, 100000),       # the same beads through PILE
    ("ccpol8sf", 512, 1, 2, 2, 12000.0, DIMER_MASS, 0.01, 100000),   # C4
    ("ccpol8sf", 1024, 1, 1, 2, 12000.0, DIMER_MASS, 0.01, 100000),  # C5
    ("1d", 64, 4, 50, 2, 10.0, [1.0], 0.05, 100000),            # C1 (fused kernel)
]


@pytest.mark.parametrize("case", BASELINE_CASES, ids=lambda c: "%s-n%d-th%d" % (c[0], c[1], c[4]))
def test_propagate_matches_oracle_at_baseline_sizes(pk, orc, case):
    """verletmodule.f90:190-250, 372-435 at the bead counts BASELINE.json quotes (C1 64, C2 256, C4 512, C5 1024), through
    the path each of them takes in production, against the oracle's literal step sequence: 1e-10 relative."""
    name, n, ntraj, steps, thermostat, beta, mass, sigma, Noutput = case
    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    a, b = _wells(name)
    vi = pk.VerletInt(pes, n, mass, beta, dt=1e-3, gamma=1.0, NMC=steps, Noutput=Noutput, seed=977).init_nm()
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, sigma, mass, seed=23)
    gid = np.arange(ntraj, dtype=np.int64) * 5 + 2
    fn = vi.propagate_pimd_pile if thermostat == 2 else vi.propagate_pimd_nm
    xg, pg, dg = fn(x, p, a, bt, dbdl, traj_gid=gid)
    for t in range(ntraj):
        orc.nm_setup(n, mass, vi.betan, 1.0, 1.0, 1e-3, False, True)
        orc.init_nm(a, bt[..., t])
        orc.set_rng(977, int(gid[t]))
        xo, po, do = orc.propagate(thermostat, x[..., t], p[..., t], dbdl[..., t], steps, 0, Noutput)
        assert relmax(xg[..., t], xo) < RTOL
        assert relmax(pg[..., t], po) < RTOL
        assert abs(dg[t] - do) <= RTOL * abs(do)


@pytest.mark.parametrize("name,n", [("2dtest", 32), ("2dtest", 160)])
def test_andersen_restart_segments_keep_the_collision_clock(pk, tmp_path, name, n):
    """restart = 1 with a FINITE Noutput under the Andersen thermostat: the reference writes its files from inside one
    loop and never touches count / rkick (verletmodule.f90:199-234), so a run cut into Noutput-step calls must resample
    momenta at the very steps of the uncut run.  (Fused kernel at n = 32, streamed path at n = 160.)"""
    pes = pk.McmodMass(name).V_init()
    a, b = _wells(name)
    ntraj, steps, nout = 4, 60, 9
    x, p, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, 0.05, [1.0])
    gid = np.arange(ntraj, dtype=np.int64) + 3
    vi = pk.VerletInt(pes, n, [1.0], 10.0, dt=1e-3, NMC=steps, imin=0, Noutput=nout, seed=5).init_nm()
    x_ref, p_ref, d_ref = vi.propagate_pimd_nm(x, p, a, bt, dbdl, traj_gid=gid)
    vs = pk.VerletInt(pes, n, [1.0], 10.0, dt=1e-3, NMC=steps, imin=0, Noutput=nout, seed=5).init_nm()
    vs.restart = 1
    xs, ps, ds = vs.propagate_restartable(1, x, p, a, bt, dbdl, traj_gid=gid, iproc=0, directory=str(tmp_path))
    assert relmax(xs, x_ref) < RTOL and relmax(ps, p_ref) < RTOL and np.abs(ds - d_ref).max() <= RTOL * np.abs(d_ref).max()
    # and the clock is NOT carried into an unrelated later call
    x2, p2, d2 = vi.propagate_pimd_nm(x, p, a, bt, dbdl, traj_gid=gid)
    assert np.array_equal(x2, x_ref) and np.array_equal(d2, d_ref)


def test_trajectory_ids_beyond_32_bits_are_rejected(pk):
    pes = pk.McmodMass("2dtest").V_init()
    a, b = _wells("2dtest")
    for n in (16, 160):   # fused and streamed
        vi = pk.VerletInt(pes, n, [1.0], 10.0, dt=1e-3, NMC=2, seed=1).init_nm()
        x, p, bt, dbdl, _ = _traj_inputs(pes, n, 2, a, b, 0.05, [1.0])
        with pytest.raises(pk.PimdkError) as ei:
            vi.propagate_pimd_pile(x, p, a, bt, dbdl, traj_gid=np.array([5, 2 ** 32], dtype=np.int64))
        assert ei.value.code == 1
        vi.propagate_pimd_pile(x, p, a, bt, dbdl, traj_gid=np.array([5, 2 ** 32 - 1], dtype=np.int64))


def test_two_live_plugins_do_not_redirect_each_other(pk, orc):
    """the library holds one PES selection and one V0; the Python mirror re-selects when another object owned it"""
    p1 = pk.McmodMass("ccpol8sf", iemonomer=0).V_init()
    p2 = pk.McmodMass("ccpol8sf", iemonomer=1).V_init()
    p3 = pk.McmodMass("2dtest").V_init()
    p3.set_V0(0.25)
    x = (GOLDEN_GEOM_ANG / 0.529177).reshape(6, 3).T
    e1, e2 = p1.V(x) * 627.510, p2.V(x) * 627.510
    assert abs(e1 - GOLDEN_VAL[2]) < 5.1e-6 and abs(e2 - GOLDEN_VALM[2]) < 3e-5
    v3 = p3.V(np.array([[3.0], [0.0]]))
    p1.V(x)
    assert p3.V(np.array([[3.0], [0.0]])) == v3 and p3.V0 == 0.25


def test_device_side_estimator_sums_match_the_host_formulas(pk):
    """pimdk_ti_reduce_dev (one rank: kernel + copy, no NCCL) against pimdk_ti_partial_sums (pimd_par.f90:397-409)"""
    import torch

    from pimd_tunneling_b200 import ti
    rng = np.random.default_rng(8)
    nintegral, nrep = 16, 37
    gid = rng.permutation(nintegral * nrep).astype(np.int64)[:500]
    dH = rng.normal(size=gid.size) * 3.0
    betan = 0.37
    ref = ti.partial_sums(dH, gid, nrep, nintegral, betan)
    d = torch.from_numpy(dH).cuda()
    gdev = torch.from_numpy(gid).cuda()
    got = ti.reduce_dev(gid.size, d.data_ptr(), gdev.data_ptr(), nrep, nintegral, betan)
    assert np.array_equal(got[:, 2], ref[:, 2])
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
    with pytest.raises(pk.PimdkError):
        ti.reduce_dev(gid.size, d.data_ptr(), gdev.data_ptr(), nrep, nintegral - 1, betan)
    assert ti.comm_info()[:2] == (0, 1)


# ---------------------------------------------------------------- opt-in analytic CCpol gradient (row N4) -----------
@pytest.mark.parametrize("nbatch", [1, 7, 1000])
def test_ccpol_analytic_gradient_mode(pk, orc, nbatch):
    """pimdk_set_mode(PIMDK_MODE_ANALYTIC): Vprime of ccpol8sf as the analytic gradient of the same energy expression.
    Parity gates: (i) against the oracle's dual-number gradient (oracle/dual.hpp), 1e-10 of max|grad| (energy 1e-12);
    (ii) against the finite-difference default at ITS truncation error, < 2e-7 of max|grad|
    (mcmod_waterdimer_ccpol.f90:40-58: eps = 1e-4 bohr; tests/test_oracle.py measures 2.6e-8 median); (iii) x untouched."""
    from pimd_tunneling_b200._lib import check, lib

    pes = pk.McmodMass("ccpol8sf").V_init()
    orc.select("ccpol8sf")
    x = thermal_dimer_geometries(nbatch, seed=31, sigma=0.06)
    v_fd, g_fd = pes.eval_batch(x)
    check(lib().pimdk_set_mode(2))
    try:
        x0 = x.copy()
        v, g = pes.eval_batch(x)
        g_only = pes.Vprime_batch(x)
        x_in = np.array(x, order="F")
        g_inplace = pes.Vprime_batch_inplace(x_in)
    finally:
        check(lib().pimdk_set_mode(0))
    assert np.array_equal(x, x0) and np.array_equal(x_in, x0)          # no finite-difference drift in this mode
    assert np.array_equal(g_only, g) and np.array_equal(g_inplace, g)
    assert np.abs(v - v_fd).max() <= 1e-12 * np.abs(v_fd).max()         # (energy-with-gradient is the analytic pipeline's own sum)
    gmax = np.abs(g_fd).max(axis=(0, 1))
    assert (np.abs(g - g_fd).max(axis=(0, 1)) / gmax).max() < 2e-7
    for k in range(min(nbatch, 24)):
        vo, go = orc.ccpol_analytic_gradient(x[:, :, k])
        assert abs(v[k] - vo) <= 1e-12 * max(1.0, abs(vo))
        assert np.abs(g[:, :, k] - go).max() <= 1e-10 * np.abs(go).max()


def test_ccpol_analytic_gradient_random_orientations(pk, orc):
    """the analytic-gradient mode over the whole range a ring polymer reaches (random relative orientations, 4.2 ... 14 bohr,
    distorted monomers; repulsive wall included): against the oracle's dual-number gradient, 1e-10 of max|grad|"""
    from pimd_tunneling_b200._lib import check, lib

    pes = pk.McmodMass("ccpol8sf").V_init()
    orc.select("ccpol8sf")
    x = random_dimer_geometries(48, seed=21)
    check(lib().pimdk_set_mode(2))
    try:
        v, g = pes.eval_batch(x)
    finally:
        check(lib().pimdk_set_mode(0))
    for k in range(x.shape[2]):
        vo, go = orc.ccpol_analytic_gradient(x[:, :, k])
        assert abs(v[k] - vo) <= 1e-12 * max(1.0, abs(vo))
        assert np.abs(g[:, :, k] - go).max() <= 1e-10 * np.abs(go).max()


def test_ccpol_analytic_mode_surfaces_and_propagation(pk, orc):
    """the mode covers the Radau-embedded potparts surfaces (3 and 10) and refuses the others; a propagation in this mode
    stays within the finite-difference truncation error of the default over a few steps"""
    from pimd_tunneling_b200._lib import check, lib

    x = thermal_dimer_geometries(5, seed=2)
    check(lib().pimdk_set_mode(2))
    try:
        orc.load_ccpol(10, 0)
        p10 = pk.McmodMass("ccpol8sf", isurf=10, iemonomer=0).V_init()
        g10 = p10.Vprime_batch(x)
        for k in range(5):
            _, go = orc.ccpol_analytic_gradient(x[:, :, k])
            assert np.abs(g10[:, :, k] - go).max() <= 1e-10 * np.abs(go).max()
        with pytest.raises(pk.PimdkError):
            pk.McmodMass("ccpol8sf", isurf=1).V_init().Vprime_batch(x)
    finally:
        orc.load_ccpol(3, 1)
        check(lib().pimdk_set_mode(0))
    pes = pk.McmodMass("ccpol8sf").V_init()
    a, b = _wells("ccpol8sf")
    n, ntraj, steps = 16, 3, 5
    vi = pk.VerletInt(pes, n, DIMER_MASS, 400.0, dt=1e-3, NMC=steps, seed=12).init_nm()
    xx, pp, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, 0.01, DIMER_MASS)
    x_fd, p_fd, d_fd = vi.propagate_pimd_pile(xx, pp, a, bt, dbdl)
    check(lib().pimdk_set_mode(2))
    try:
        x_an, p_an, d_an = vi.propagate_pimd_pile(xx, pp, a, bt, dbdl)
    finally:
        check(lib().pimdk_set_mode(0))
    assert relmax(x_an, x_fd) < 1e-9 and relmax(p_an, p_fd) < 1e-6 and np.abs(d_an - d_fd).max() <= 1e-8 * np.abs(d_fd).max()


# ---------------------------------------------------------------- a further in-repo surface: mcmod_so2 (row N4) -----
def test_so2_surface_bit_exact_and_propagation(pk, orc):
    """mcmod_so2.f90 (harmonic ring, analytic gradient, closed-form Hessian) behind pimdk_pes_select("so2"): V, Vprime and
    Vdoubleprime bit-exact against the oracle; propagation (fused kernel at n = 16, streamed path at n = 160, both
    thermostats) within 1e-10."""
    orc.select("so2")
    rng = np.random.default_rng(7)
    for params in (None, [3.0, 2.5]):
        pes = pk.McmodMass("so2", params=params).V_init()
        om, r0 = (10000.0, 20.0) if params is None else params
        orc.set_so2(om, r0)
        x = np.asfortranarray(rng.normal(size=(2, 1, 513)) * (8.0 if params is None else 1.5) + 1.0)
        v, g = pes.eval_batch(x)
        vo, go, _ = orc.pes_eval(x)
        assert np.array_equal(v, vo) and np.array_equal(g, go)
        h = pes.Vdoubleprime_batch(np.array(x[:, :, :40], order="F"))
        for k in range(40):
            ho = np.empty((2, 1, 2, 1), order="F")
            xk = np.array(x[:, :, k], order="F")
            orc.L.orc_Vdoubleprime(xk.ctypes.data_as(ctypes_P), ho.ctypes.data_as(ctypes_P))
            assert np.array_equal(h[..., k], ho)
    pes = pk.McmodMass("so2", params=[3.0, 2.5]).V_init()
    orc.set_so2(3.0, 2.5)
    a = np.asfortranarray(np.array([[2.5], [0.0]]))
    b = np.asfortranarray(np.array([[2.5 * np.cos(1.0)], [2.5 * np.sin(1.0)]]))
    for n, thermostat, nout in ((16, 2, 100000), (16, 1, 6), (160, 2, 100000), (160, 1, 9)):
        vi = pk.VerletInt(pes, n, [1.0], 10.0, dt=1e-3, NMC=30, Noutput=nout, seed=3).init_nm()
        x, p, bt, dbdl, _ = _traj_inputs(pes, n, 3, a, b, 0.05, [1.0])
        gid = np.arange(3, dtype=np.int64) + 40
        fn = vi.propagate_pimd_pile if thermostat == 2 else vi.propagate_pimd_nm
        xg, pg, dg = fn(x, p, a, bt, dbdl, traj_gid=gid)
        for t in range(3):
            orc.nm_setup(n, [1.0], vi.betan, 1.0, 1.0, 1e-3, False, True)
            orc.init_nm(a, bt[..., t])
            orc.set_rng(3, int(gid[t]))
            xo, po, do = orc.propagate(thermostat, x[..., t], p[..., t], dbdl[..., t], 30, 0, nout)
            assert relmax(xg[..., t], xo) < RTOL and relmax(pg[..., t], po) < RTOL and abs(dg[t] - do) <= RTOL * abs(do)
    orc.select("ccpol8sf")


@pytest.mark.parametrize("name,n", [("2dtest", 16), ("2dtest", 160), ("1d", 24)])
def test_dhdrlimit_outlier_reinitialisation(pk, orc, name, n):
    """dHdrlimit (pimd_par.f90:45,88; verletmodule.f90:404-409): an over-limit estimator contribution is dropped and the ring
    polymer re-initialised by init_path.  The limit is set inside the range the contribution actually takes, so that some
    trajectories trip it several times; fused kernel (n <= 128) and streamed path against the oracle, 1e-10; and the
    Andersen thermostat ignores the limit like propagate_pimd_nm."""
    from pimd_tunneling_b200 import path as P

    mass = [1.0]
    pes = pk.McmodMass(name).V_init()
    orc.select(name)
    a, b = _wells(name)
    ntraj, steps = 5, 40
    vi = pk.VerletInt(pes, n, mass, 10.0, dt=1e-3, NMC=steps, seed=808).init_nm()
    x, p, bt, dbdl, (lam, path, spl, xi) = _traj_inputs(pes, n, ntraj, a, b, 0.05, mass)
    gid = np.arange(ntraj, dtype=np.int64) + 1
    x0, p0 = vi.init_path(xi, lam, path, spl, traj_gid=gid)
    x_ref, p_ref, d_ref = vi.propagate_pimd_pile(x0, p0, a, bt, dbdl, traj_gid=gid)
    # contribution of the initial state: a limit just above the smallest |contr| trips for the others
    contr0 = np.abs(np.einsum("k,jkt,jkt->t", np.asarray(mass), -x0[n - 1], dbdl))
    limit = float(np.sort(contr0)[1] * 1.0001)
    vi.set_dhdrlimit(limit, xi, lam, path, spl)
    try:
        xg, pg, dg = vi.propagate_pimd_pile(x0, p0, a, bt, dbdl, traj_gid=gid)
        xa, pa, da = vi.propagate_pimd_nm(x0, p0, a, bt, dbdl, traj_gid=gid)
    finally:
        vi.set_dhdrlimit(-1.0)
    xb, pb, db = vi.propagate_pimd_nm(x0, p0, a, bt, dbdl, traj_gid=gid)
    assert np.array_equal(xa, xb) and np.array_equal(da, db)          # Andersen: no guard
    assert not np.array_equal(dg, d_ref)                                # the guard did act
    tripped = 0
    for t in range(ntraj):
        orc.nm_setup(n, mass, vi.betan, 1.0, 1.0, 1e-3, False, True)
        orc.init_nm(a, bt[..., t])
        orc.set_rng(808, int(gid[t]))
        orc.set_dhdrlimit(limit, float(xi[t]), lam, path, spl)
        xo, po, do = orc.propagate(2, x0[..., t], p0[..., t], dbdl[..., t], steps, 0, 100000)
        assert relmax(xg[..., t], xo) < RTOL and relmax(pg[..., t], po) < RTOL
        assert abs(dg[t] - do) <= RTOL * max(abs(do), np.abs(d_ref).max())
        tripped += int(abs(dg[t] - d_ref[t]) > 1e-12 * abs(d_ref[t]))
    assert 1 <= tripped <= ntraj


def test_water_methane_surface_bit_exact(pk, orc):
    """mcmod_watmeth.f90 + watermethane.f90 (wmrb, wmrb_grad: 63 Tang-Toennies site pairs with Numerical Recipes' incomplete
    gamma function) behind pimdk_pes_select("watmeth"): V, Vprime, Vdoubleprime bit-exact against the oracle on rigid-body
    geometries (ragged batch), and two PILE steps of a small ring polymer within 1e-10."""
    from oracle_lib import watmeth_geometries

    pes = pk.McmodMass("watmeth").V_init()
    orc.select("watmeth")
    x = watmeth_geometries(301, seed=9)
    v, g = pes.eval_batch(x)
    vo, go, _ = orc.pes_eval(x)
    assert np.array_equal(v, vo) and np.array_equal(g, go)
    pes.set_V0(0.5)                                   # mcmod_watmeth.f90:15-27: V does not subtract V0
    assert np.array_equal(pes.V_batch(x[:, :, :3]), vo[:3])
    xs = np.array(x[:, :, :2], order="F")
    h = pes.Vdoubleprime_batch(xs)
    for k in range(2):
        xk = np.array(x[:, :, k], order="F")
        ho = np.empty((3, 17, 3, 17), order="F")
        orc.L.orc_Vdoubleprime(xk.ctypes.data_as(ctypes_P), ho.ctypes.data_as(ctypes_P))
        assert np.array_equal(h[..., k], ho) and np.array_equal(xs[:, :, k], xk)
    mass = [1837.0] * 17
    a, b = np.asfortranarray(x[:, :, 0]), np.asfortranarray(x[:, :, 1])
    n, ntraj, steps = 8, 2, 2
    vi = pk.VerletInt(pes, n, mass, 800.0, dt=1e-3, NMC=steps, seed=4).init_nm()
    xx, pp, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, 0.002, mass)
    xg, pg, dg = vi.propagate_pimd_pile(xx, pp, a, bt, dbdl)
    for t in range(ntraj):
        orc.nm_setup(n, mass, vi.betan, 1.0, 1.0, 1e-3, False, True)
        orc.init_nm(a, bt[..., t])
        orc.set_rng(4, t)
        xo, po, do = orc.propagate(2, xx[..., t], pp[..., t], dbdl[..., t], steps, 0, 100000)
        assert relmax(xg[..., t], xo) < RTOL and relmax(pg[..., t], po) < RTOL and abs(dg[t] - do) <= RTOL * abs(do)
    orc.select("ccpol8sf")


def test_malonaldehyde_surface_bit_exact(pk, orc):
    """mcmod_malon.f90 + pes_malonaldehyde.f90 (`pes`: 9 Morse terms and 3549 Gaussians over the 36 distances, analytic gradient
    through the B matrix, analytic Hessian) behind pimdk_pes_select("malon"): V, Vprime, Vdoubleprime bit-exact against the
    oracle (ragged batch, V0 subtracted), the reference's own minimum-energy structure gives V = 0, and two PILE steps of a
    small ring polymer (27 degrees of freedom per bead, streamed path) stay within 1e-10."""
    from oracle_lib import MALON_BOHR, MALON_MASS, MALON_MIN_ANG, malon_geometries

    pes = pk.McmodMass("malon").V_init()
    orc.select("malon")
    x = malon_geometries(261, seed=7)
    v, g = pes.eval_batch(x)
    vo, go, _ = orc.pes_eval(x)
    assert np.array_equal(v, vo) and np.array_equal(g, go)
    xmin = np.asfortranarray((MALON_MIN_ANG / MALON_BOHR).T.reshape(3, 9, 1))
    assert abs(pes.V_batch(xmin)[0]) < 1e-12                       # pes_malonaldehyde.f90:12-21: energy above equilibrium
    pes.set_V0(0.125)                                              # mcmod_malon.f90:21
    orc.set_V0(0.125)
    assert np.array_equal(pes.V_batch(x[:, :, :5]), orc.pes_eval(x[:, :, :5], gradient=False)[0])
    pes.set_V0(0.0)
    orc.set_V0(0.0)
    xs = np.array(x[:, :, :3], order="F")
    h = pes.Vdoubleprime_batch(xs)
    for k in range(3):
        ho, _ = orc.Vdoubleprime(x[:, :, k])
        assert np.array_equal(h[..., k], ho) and np.array_equal(xs[:, :, k], x[:, :, k])
    mass = list(MALON_MASS)
    a, b = np.asfortranarray(x[:, :, 0]), np.asfortranarray(x[:, :, 1])
    n, ntraj, steps = 8, 2, 2
    vi = pk.VerletInt(pes, n, mass, 800.0, dt=1e-3, NMC=steps, seed=4).init_nm()
    xx, pp, bt, dbdl, _ = _traj_inputs(pes, n, ntraj, a, b, 0.002, mass)
    xg, pg, dg = vi.propagate_pimd_pile(xx, pp, a, bt, dbdl)
    for t in range(ntraj):
        orc.nm_setup(n, mass, vi.betan, 1.0, 1.0, 1e-3, False, True)
        orc.init_nm(a, bt[..., t])
        orc.set_rng(4, t)
        xo, po, do = orc.propagate(2, xx[..., t], pp[..., t], dbdl[..., t], steps, 0, 100000)
        assert relmax(xg[..., t], xo) < RTOL and relmax(pg[..., t], po) < RTOL and abs(dg[t] - do) <= RTOL * abs(do)
    orc.select("ccpol8sf")
