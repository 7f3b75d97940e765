#!/usr/bin/env python3
"""Drive the streamed propagation path on a BASELINE-shaped state for profiling (run it under ncu with a kernel filter,
or bare: it then prints the per-family CUDA-event times of the library's own profiling spans).

    python tools/prof_path.py --pes ccpol8sf --n 512 --ntraj 2048 --thermostat 2 --steps 2 [--noutput 1] [--mode 0]

ntraj = 2048 ring polymers of 512 beads x 18 dof is 151 MB per state array: larger than the 126 MB L2, so the
elementwise kernels see HBM like they do in the full C4 run."""
import argparse
import ctypes
import os
import sys

import numpy as np

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch  # noqa: E402

import pimd_tunneling_b200 as pk  # noqa: E402
from bench import ti_path, wells  # noqa: E402
from pimd_tunneling_b200 import path as P  # noqa: E402
from pimd_tunneling_b200._lib import check, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pes", default="ccpol8sf")
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--ntraj", type=int, default=2048)
ap.add_argument("--thermostat", type=int, default=2)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--noutput", type=int, default=100000)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--beta", type=float, default=0.0)
ap.add_argument("--gemm", type=int, default=1)
args = ap.parse_args()

pk.init(0)
L = lib()
check(L.pimdk_set_mode(args.mode))
check(L.pimdk_set_gemm(args.gemm))
pes = pk.McmodMass(args.pes).V_init()
a, b, mass = wells(args.pes)
beta = args.beta or (12000.0 if args.pes == "ccpol8sf" else 10.0)
vi = pk.VerletInt(pes, args.n, mass, beta, dt=1e-3, NMC=1, Noutput=args.noutput, seed=99).init_nm()
lam, path, spl = ti_path(args.pes, a, b)
n, ndof, ntraj = args.n, pes.ndof, args.ntraj
xi = np.linspace(0.1, 0.9, ntraj)
bt, dbdl = P.endpoints(lam, path, spl, xi)
dev = torch.device("cuda", 0)
x_d = torch.empty((ntraj, ndof, n), dtype=torch.float64, device=dev)
p_d = torch.empty_like(x_d)
chunk = max(1, (1 << 26) // (n * ndof))
for lo in range(0, ntraj, chunk):
    hi = min(ntraj, lo + chunk)
    xc, pc = vi.init_path(xi[lo:hi], lam, path, spl, traj_gid=np.arange(lo, hi, dtype=np.int64))
    x_d[lo:hi] = torch.from_numpy(np.ascontiguousarray(xc.reshape(-1, order="F").reshape(hi - lo, ndof, n))).to(dev)
    p_d[lo:hi] = torch.from_numpy(np.ascontiguousarray(pc.reshape(-1, order="F").reshape(hi - lo, ndof, n))).to(dev)
a_d = torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).to(dev)
b_d = torch.from_numpy(np.ascontiguousarray(bt.reshape(-1, order="F"))).to(dev)
dbdl_d = torch.from_numpy(np.ascontiguousarray(dbdl.reshape(-1, order="F"))).to(dev)
dH = torch.zeros(ntraj, dtype=torch.float64, device=dev)


def call(k):
    vi.propagate_dev(args.thermostat, ntraj, x_d.data_ptr(), p_d.data_ptr(), a_d.data_ptr(), b_d.data_ptr(), dbdl_d.data_ptr(),
                     None, dH.data_ptr(), NMC=k)


if args.warm:
    call(args.warm)
torch.cuda.synchronize()
check(L.pimdk_profile(1))
check(L.pimdk_profile_reset())
call(args.steps)
torch.cuda.synchronize()
out = {}
for f in ("pes", "gemm", "update", "estimator", "fused"):
    m_, c_ = ctypes.c_double(), ctypes.c_int64()
    check(L.pimdk_profile_get(f.encode(), ctypes.byref(m_), ctypes.byref(c_)))
    out[f] = (round(m_.value, 4), int(c_.value))
rows = ntraj * ndof
gl = out["gemm"][1]
print("families (ms, launches):", out)
if gl:
    print("transform: %.4f ms per launch, %.2f TFLOP/s (rows %d, n %d, engine %d)" % (out["gemm"][0] / gl, 2.0 * rows * n * n * gl / (out["gemm"][0] * 1e-3) / 1e12, rows, n, args.gemm))
el = rows * n
if out["update"][1]:
    print("update family: %.4f ms per launch; one nm_update2 moves %.3f GB (P,Q rw + G r + QB w + BV r)" % (out["update"][0] / out["update"][1], el * 8 * 7 / 1e9))
pk.finalize()
