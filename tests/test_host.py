"""CPU tests of the host side: the C ABI library loads and exports every symbol include/pimdk.h
declares, compute calls fail loudly without a GPU, the path/TI bookkeeping matches the reference's
formulas, and the multi-rank reduction works over gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from pimd_tunneling_b200._lib import LIB_PATH, lib

    hdr = open(os.path.join(ROOT, "include", "pimdk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pimdk_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libpimdk.so does not export %s" % name
    out = subprocess.run(["nm", "-D", LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (pimdk_[a-z0-9_]+)", out))
    assert declared <= exported


def test_product_does_not_touch_the_oracle():
    """the shipped path must not import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "pimd_tunneling_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r'#include\s+"[^"]*oracle/', text), (dp, f)
                assert "liboracle" not in text and "oracle_lib" not in text, (dp, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libpimdk.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


@pytest.mark.skipif(_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    import pimd_tunneling_b200 as pk

    with pytest.raises(pk.PimdkError) as ei:
        pk.init()
    assert ei.value.code == 2 and "no CPU path" in str(ei.value)
    pes = pk.McmodMass("1d")
    with pytest.raises(Exception):
        pes.V_init()
    with pytest.raises(RuntimeError):
        pes.V(np.zeros((1, 1)))


def test_gauleg_and_ti_statistics_match_reference_formulas():
    from pimd_tunneling_b200 import VerletInt, ti

    x, w = VerletInt.gauleg(0.0, 1.0, 16)
    xr, wr = np.polynomial.legendre.leggauss(16)
    assert np.abs(np.sort(x) - (xr + 1) / 2).max() < 1e-13 and abs(w.sum() - 1.0) < 1e-13
    nint, nrep, betan = 16, 8, 0.31
    rng = np.random.default_rng(0)
    dH = rng.normal(1.0, 0.3, nint * nrep)
    gid = ti.global_ids(nint, nrep)
    sums = ti.partial_sums(dH, gid, nrep, nint, betan)
    res = ti.finish(sums, w, betan)
    # pimd_par.f90:401-424 restated
    I = (dH / betan ** 2).reshape(nint, nrep)
    mean = I.mean(axis=1)
    var = (I ** 2).mean(axis=1) - mean ** 2
    assert np.allclose(res["mean"], mean, rtol=1e-13) and np.allclose(res["var"], var, rtol=1e-9, atol=1e-12)
    assert abs(res["deltaA"] - np.sum(w * mean)) < 1e-12 * abs(res["deltaA"])
    assert abs(res["sigmaA"] - np.sqrt(np.sum(w ** 2 * var))) < 1e-10
    assert abs(res["q_over_q0"] - np.exp(-res["deltaA"] * betan)) < 1e-15
    # block partition of pimd_par.f90:109-110
    assert [ti.shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sum(hi - lo for lo, hi in (ti.shard(8192, r, 8) for r in range(8))) == 8192


def test_path_module_matches_scipy():
    from scipy.interpolate import CubicSpline

    from pimd_tunneling_b200 import path as P

    pts = np.zeros((7, 2, 1), order="F")
    t = np.linspace(0, np.pi / 3, 7)
    pts[:, 0, 0], pts[:, 1, 0] = 3 * np.cos(t), 3 * np.sin(t)
    lam, path, spl = P.build_path(pts)
    assert lam[0] == 0.0 and lam[-1] == 1.0 and np.all(np.diff(lam) > 0)
    cs = CubicSpline(lam, path[:, 1, 0], bc_type="natural")
    xint, dbd = P.endpoints(lam, path, spl, [0.0, 0.3, 0.77, 1.0])
    for k, x in enumerate([0.0, 0.3, 0.77, 1.0]):
        assert abs(xint[1, 0, k] - cs(x)) < 1e-13 and abs(dbd[1, 0, k] - cs(x, 1)) < 1e-12
    # straight line: spline of a line is the line
    lam, path, spl = P.build_path(np.array([[[-1.0]], [[1.0]]]))
    xint, dbd = P.endpoints(lam, path, spl, [0.25])
    assert xint[0, 0, 0] == -0.5 and dbd[0, 0, 0] == 2.0


_GLOO_WORKER = r"""
import os, sys, numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from pimd_tunneling_b200 import ti
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
nint, nrep, betan = 8, 6, 0.4
rng = np.random.default_rng(5)
dH = rng.normal(2.0, 0.5, nint * nrep)
gid = ti.global_ids(nint, nrep)
lo, hi = ti.shard(nint * nrep, dist.get_rank(), 2)
sums = ti.allreduce_sums(ti.partial_sums(dH[lo:hi], gid[lo:hi], nrep, nint, betan))
full = ti.partial_sums(dH, gid, nrep, nint, betan)
assert np.allclose(sums, full, rtol=1e-14, atol=0), (sums, full)
w = np.full(nint, 1.0 / nint)
r = ti.finish(sums, w, betan)
assert abs(r["deltaA"] - ti.finish(full, w, betan)["deltaA"]) < 1e-13
dist.destroy_process_group()
print("rank", sys.argv[1], "ok")
"""


def test_two_rank_estimator_allreduce_gloo(tmp_path):
    """the N>1 path: shard trajectories by global id, ONE all-reduce of {sum, sum^2, count} per lambda"""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT, "port": 29000 + os.getpid() % 2000})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_namelist_reader_defaults_and_overrides():
    from pimd_tunneling_b200.ti_driver import MCData, read_namelist

    d = MCData()
    assert (d.n, d.beta, d.NMC, d.Noutput, d.dt, d.nintegral, d.nrep, d.thermostat) == (100, 100.0, 5000000, 100000, 1e-3, 5, 1, 1)
    mc = read_namelist("&MCDATA\n n=64, beta=10.0d0, NMC=2000, thermostat=2,\n nintegral=8, nrep=4, cayley=.true., basename='x'\n/\n")
    assert (mc.n, mc.beta, mc.NMC, mc.thermostat, mc.nintegral, mc.nrep, mc.cayley) == (64, 10.0, 2000, 2, 8, 4, True)
    assert mc.extra == {"basename": "x"} and mc.tau == 1.0 and mc.gamma == 1.0


def test_multiwell_waterdimer_postprocessing():
    """row N4: the two-temperature projection of python_utils/multiwell_waterdimer.py on its own input numbers"""
    import numpy as np
    from pimd_tunneling_b200 import multiwell as mw

    rho1 = mw.class_weights(2.0 * 0.111478983128910, 0.0, 1.071538742584852e-3, 5.632020635958411e-3, 0.0)
    rho2 = mw.class_weights(2.0 * 0.230546268494811, 0.0, 3.118506023220292e-3, 1.049548502801190e-2, 0.0)
    out = mw.waterdimer_levels(rho1, rho2, 12000.0, 20000.0)
    assert list(out) == ["A1+", "E+", "B1+", "A2-", "E-", "B2-"]
    assert out["A1+"]["level_cm"] == 0.0 and out["A1+"]["I1"] == 0.0        # totally symmetric state is the origin
    # independent evaluation of one level: E- with characters (1,-1,1,0,0,0,0,-1)
    c = np.array([1, -1, 1, 0, 0, 0, 0, -1.0])
    i1 = ((1 - c) * rho1).sum() / ((1 + c) * rho1).sum()
    i2 = ((1 - c) * rho2).sum() / ((1 + c) * rho2).sum()
    ref = 219475.0 * 2.0 * (np.arctanh(i2) - np.arctanh(i1)) / 8000.0
    assert abs(out["E-"]["level_cm"] - ref) <= 1e-12 * abs(ref)
    lv = [out[k]["level_cm"] for k in out]
    assert all(np.isfinite(lv)) and lv[3] > lv[0] and lv[4] > 0      # acceptor-tunnelling partners lie above A1+


def test_alignment_of_the_wells():
    """get_align / align_atoms / rotate_atoms (instantonmod.f90:222-376) as pimd_par.f90:159-165 uses them: atom1 to the
    origin, atom1->atom2 on the x axis, atom3 in the xz plane; rigid motion (distances kept); well2 carried by well1's
    angles unless alignwell; identity below the reference's 1e-10 angle threshold."""
    from pimd_tunneling_b200 import path as P

    rng = np.random.default_rng(3)
    w1, w2 = rng.normal(size=(3, 6)), rng.normal(size=(3, 6))
    a1, a2 = P.align_wells(w1, w2)
    assert np.abs(a1[:, 0]).max() == 0.0 and np.abs(a1[1:, 1]).max() < 1e-15 and abs(a1[1, 2]) < 1e-15 and a1[0, 1] > 0
    dist = lambda a: np.linalg.norm(a[:, :, None] - a[:, None, :], axis=0)
    assert np.abs(dist(a1) - dist(w1)).max() < 1e-14 and np.abs(dist(a2) - dist(w2)).max() < 1e-14
    # well2 moved by well1's rotation: relative orientation of the two wells is kept
    t = P.get_align(w1)
    r = lambda x: P.rotate_atoms(P.rotate_atoms(P.rotate_atoms(x, 3, t[0]), 2, t[1]), 1, t[2])
    assert np.abs(a2 - r(w2 - w2[:, :1])).max() < 1e-14
    b1, b2 = P.align_wells(w1, w2, alignwell=True)
    assert np.array_equal(b1, a1) and np.abs(b2[1:, 1]).max() < 1e-15 and abs(b2[1, 2]) < 1e-15
    # already aligned: all three angles below 1e-10 -> the reference copies the input (no shift to atom1 either)
    assert np.array_equal(P.align_atoms(a1 + 5.0, 0.0, 0.0, 0.0), a1 + 5.0)
    with pytest.raises(ValueError):
        P.get_align(np.zeros((2, 3)))


def test_read_path_pieces(tmp_path):
    """read_path (instantonmod.f90:895-1035) on the host: xyz frames (xunit = 2 converts from Angstrom), arc-length
    lampath, findmiddle's bisection for the barrier top, the `centre` reparametrisation that moves it to 1/2, and the
    instanton refinement hook (path = well1, xtilde, well2)."""
    from pimd_tunneling_b200 import path as P

    xs = np.array([-1.0, -0.8, -0.55, -0.2, 0.25, 0.5, 0.7, 0.9, 1.0])
    f = tmp_path / "path.xyz"
    f.write_text("".join("1\nframe %d\nX %.10f\n" % (i, x * 0.529177) for i, x in enumerate(xs)))
    pts = P.read_xyz_frames(str(f), 1, 1, xunit=2)
    assert pts.shape == (9, 1, 1) and np.abs(pts[:, 0, 0] - xs).max() < 1e-9
    V = lambda x: (x[0, 0, :] ** 2 - 1.0) ** 2
    r = P.read_path(pts, V, n=7)
    assert np.abs(r["lampath"] - (xs + 1.0) / 2.0).max() < 1e-9 and np.abs(r["Vpath"] - (xs ** 2 - 1) ** 2).max() < 1e-9
    assert np.abs(r["xtilde"][:, 0, 0] - np.linspace(-1, 1, 7)).max() < 1e-3
    xm = P.findmiddle(0.3, 0.7, r["lampath"], r["Vpath"])
    assert abs(xm - 0.5) < 2e-2                         # top of the splined barrier: x = 0 <-> lambda = 1/2
    # an asymmetric parametrisation is pulled back to the middle
    lam2 = r["lampath"] ** 1.6
    lam2 /= lam2[-1]
    xm2 = P.findmiddle(0.2, 0.8, lam2, r["Vpath"])
    lamc, a, b, xmid = P.centre_lampath(lam2, r["Vpath"]) if 0.3 < xm2 < 0.7 else (None, 0, 0, xm2)
    if lamc is not None:
        assert abs(a * 0.25 + b * 0.5 - xmid) < 1e-12      # the quadratic map sends 1/2 to the old barrier position
        assert lamc[0] == 0.0 and abs(lamc[-1] - 1.0) < 1e-12 and np.all(np.diff(lamc) > 0)
    # instanton hook with fixed ends: new path = well1, refined beads, well2
    w1, w2 = np.array([[-1.0]]), np.array([[1.0]])
    r2 = P.read_path(pts, V, n=7, instanton=lambda xt: xt * 0.9, well1=w1, well2=w2)
    assert r2["path"].shape == (9, 1, 1) and r2["path"][0, 0, 0] == -1.0 and r2["path"][-1, 0, 0] == 1.0
    assert np.abs(r2["path"][1:-1, 0, 0] - 0.9 * r["xtilde"][:, 0, 0]).max() < 1e-15


def _c_prototypes():
    """{name: (return kind, [argument kinds])} of include/pimdk.h; kinds: i64 / f64 by value, ptr"""
    hdr = open(os.path.join(ROOT, "include", "pimdk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"^((?:const\s+)?[A-Za-z_0-9]+\s*\*?)\s*(pimdk_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr, flags=re.M):
        kinds = []
        for a in [s.strip() for s in args.replace("\n", " ").split(",")]:
            if a in ("void", ""):
                continue
            if "*" in a:
                kinds.append("ptr")
            elif re.match(r"(pimdk_int|uint64_t|int64_t)\b", a):
                kinds.append("i64")
            elif re.match(r"double\b", a):
                kinds.append("f64")
            else:
                raise AssertionError("unclassified C argument %r of %s" % (a, name))
        r = ret.replace(" ", "")
        protos[name] = ({"int": "int", "pimdk_int": "i64", "constchar*": "ptr"}[r], kinds)
    return protos


def _fortran_interfaces():
    """the same from the interface block of fortran/pimdk_mod.f90 (no Fortran compiler exists in this image, so the
    shim cannot be compiled here: this is the check that it at least agrees with the header argument by argument)"""
    src = open(os.path.join(ROOT, "fortran", "pimdk_mod.f90")).read()
    block = src[src.index("interface"):src.index("end interface")]
    block = re.sub(r"&\s*\n\s*", " ", block)
    out = {}
    for m in re.finditer(r"^\s*(integer\(c_int\)|integer\(c_int64_t\)|type\(c_ptr\))\s+function\s+(\w+)\s*\(([^)]*)\)\s*"
                         r"bind\(C,\s*name=\"(\w+)\"\)(.*?)end function", block, flags=re.S | re.M | re.I):
        ret, fname, dummies, cname, body = m.groups()
        assert fname == cname
        dummies = [d.strip().lower() for d in dummies.split(",") if d.strip()]
        kind = {}
        body = "\n".join(l.split("!")[0] for l in body.splitlines())
        for stmt in re.split(r"[;\n]", body):
            stmt = stmt.strip()
            if "::" not in stmt:
                continue
            decl, names = stmt.split("::")
            decl = decl.lower()
            byval = "value" in decl
            for nm in re.split(r",\s*(?![^()]*\))", names):
                nm = nm.strip().lower()
                arr = nm.endswith("(*)")
                nm = nm.replace("(*)", "")
                if decl.startswith("type(c_ptr)"):
                    assert byval and not arr, (fname, nm)
                    kind[nm] = "ptr"
                elif decl.startswith("integer(c_int64_t)"):
                    kind[nm] = "i64" if byval else "ptr"
                    assert byval != arr, (fname, nm)
                elif decl.startswith("real(c_double)"):
                    kind[nm] = "f64" if byval else "ptr"
                    assert byval != arr, (fname, nm)   # a non-value scalar would still be a pointer, but none is meant
                elif decl.startswith("character(kind=c_char)"):
                    assert arr and not byval, (fname, nm)
                    kind[nm] = "ptr"
                else:
                    raise AssertionError("unclassified Fortran declaration %r in %s" % (stmt, fname))
        assert set(kind) == set(dummies), (fname, sorted(kind), dummies)
        r = {"integer(c_int)": "int", "integer(c_int64_t)": "i64", "type(c_ptr)": "ptr"}[ret.lower()]
        out[cname] = (r, [kind[d] for d in dummies])
    return out


def test_fortran_shim_agrees_with_the_c_header():
    c = _c_prototypes()
    f = _fortran_interfaces()
    assert len(f) >= 22 and len(c) >= 40
    for name, sig in f.items():
        assert name in c, "%s bound in fortran/pimdk_mod.f90 but not declared in include/pimdk.h" % name
        assert sig == c[name], (name, sig, c[name])
    # every entry point the shim's executable part calls is in its interface block
    src = open(os.path.join(ROOT, "fortran", "pimdk_mod.f90")).read()
    body = src[src.index("end interface"):]
    body = "\n".join(l.split("!")[0] for l in body.splitlines())
    own = set(re.findall(r"^\s*subroutine\s+(pimdk_[a-z0-9_]+)", body, flags=re.M | re.I))   # the shim's own Fortran procedures
    assert {"pimdk_check", "pimdk_propagate_tasks", "pimdk_ti_statistics", "pimdk_write_restart"} <= own
    called = set(re.findall(r"\b(pimdk_[a-z0-9_]+)\s*\(", body)) - own
    assert called and called <= set(f), called - set(f)


def test_header_is_plain_c(tmp_path):
    """include/pimdk.h is the drop-in boundary: it must compile as C99 (no C++, no torch types) and a C program
    must link against libpimdk.so through it"""
    from pimd_tunneling_b200._lib import LIB_PATH

    src = tmp_path / "t.c"
    src.write_text('#include "pimdk.h"\n#include <stdio.h>\nint main(void) {\n'
                   '  double x[2], w[2];\n  int rc = pimdk_gauleg(0.0, 1.0, 2, x, w);\n'
                   '  printf("%d %.17g %.17g\\n", rc, x[0] + x[1], w[0] + w[1]);\n'
                   '  return pimdk_last_error() == 0;\n}\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), LIB_PATH, "-Wl,-rpath," + os.path.dirname(LIB_PATH)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "0" and abs(float(out[1]) - 1.0) < 1e-15 and abs(float(out[2]) - 1.0) < 1e-15


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract: stdout carries ONE JSON line (library banners go to stderr: file descriptor 1 is pointed at
    stderr for the run).  The reference arm needs no GPU: it times the CPU oracle on a bounded sample."""
    import json

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "bead-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("C4:")
