#!/usr/bin/env python3
"""bench.py — ring-polymer bead-steps/s of the hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c4|c5|c3|c2|c1]
    torchrun ... bench.py --gpus N ...      (one rank per GPU, NCCL)

A "step" is one thermostatted Verlet step (PES gradient on every bead + normal-mode transforms +
thermostat + TI estimator) of ALL ring polymers resident on the GPU.  Default workload = BASELINE
config C4: CCpol-8sf water dimer, 512 beads x 8192 trajectories (16 lambda x 512 rep), PILE thermostat,
beta = 12000 a.u., dt = 1e-3; per-GPU work is fixed as N grows (weak scaling: trajectories are
independent units; the only exchange is one all-reduce of 3*nintegral estimator sums per call).

  value  : device-resident throughput (state in HBM before the timed region), CUDA events, max over ranks
  e2e    : same metric through the host-buffer C ABI call pimdk_propagate (H2D + D2H inside the region)
  roofline: dominant kernel = CCpol finite-difference gradient; achieved = algorithmic FP64 flop
           (exact source-level census of the reference arithmetic) / measured kernel time;
           peak = DFMA pipe measured live by pimdk_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline / --impl reference: the CPU oracle (C++ restatement of the reference's algorithm; the
           Fortran reference cannot be compiled in this image) on all host cores, bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_O, M_H = 15.9949146221 * 1822.888486, 1.0078250321 * 1822.888486

# exact source-level FP64 operation census of ONE `ccpol` energy at the known-answer geometry
# (oracle/opcount.hpp, tests/test_oracle.py::test_opcount): add 25156, mul 33156, div 1760, sqrt 806,
# exp 1129, pow 33, trig 28
FLOP_PER_ENERGY = 25156 + 33156 + 1760 + 806 + 1129 + 33 + 28
FLOP_PER_BEAD_GRAD = {"ccpol8sf": 36 * FLOP_PER_ENERGY + 36, "2dtest": 6 * (2 + 14) + 12, "1d": 8}
# Counters that only a profiler can give (DRAM bytes and executed FP64 flop per bead-gradient of the CCpol pipeline)
# are NOT pasted here: they are read from the committed summary of the `ncu --set full` capture
# (profiles/ccpol_counters.json, written by tools/ncu_pipeline_summary.py) together with the hash of the kernel
# sources the capture was taken on; a summary whose hash differs from the tree's is reported as stale and not used.
CCPOL_SOURCES = {
    "fd": ["pimd_tunneling_b200/csrc/ccpol_kernels.cu", "pimd_tunneling_b200/csrc/ccpol_device.cuh",
           "pimd_tunneling_b200/csrc/ccpol_tables.h", "include/pimdk_detmath.h"],
    "analytic": ["pimd_tunneling_b200/csrc/ccpol_grad_kernels.cu", "pimd_tunneling_b200/csrc/ccpol_grad.cuh",
                 "pimd_tunneling_b200/csrc/ccpol_tables.h", "include/pimdk_detmath.h"],
}


def ccpol_source_hash(mode="strict"):
    import hashlib

    h = hashlib.sha256()
    for f in CCPOL_SOURCES["analytic" if mode == "analytic" else "fd"]:
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def ccpol_counters(mode):
    """{"dram_bytes_per_bead", "sass_flop_per_bead", "source"} for `mode`, or None with the reason"""
    path = os.path.join(ROOT, "profiles", "ccpol_counters.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except Exception:
        return None, "profiles/ccpol_counters.json absent"
    ent = d.get(mode)
    if not ent:
        return None, "no ncu capture of mode %s in profiles/ccpol_counters.json" % mode
    if ent.get("source_sha256") != ccpol_source_hash(mode):
        return None, "stale: %s was captured on other kernel sources (%s...)" % (ent.get("capture"), str(ent.get("source_sha256"))[:12])
    return ent, "profiles/ccpol_counters.json <- %s" % ent.get("capture")


CONFIGS = {
    # name: pes, n, nintegral, nrep, thermostat, beta, Noutput
    "c4": dict(pes="ccpol8sf", n=512, nintegral=16, nrep=512, thermostat=2, beta=12000.0, Noutput=100000,
               label="C4: CCpol-8sf water dimer acceptor-switch TI, 512 beads x 8192 trajectories, 16 lambda, PILE"),
    "c5": dict(pes="ccpol8sf", n=1024, nintegral=16, nrep=256, thermostat=2, beta=12000.0, Noutput=100000,
               label="C5: CCpol-8sf water dimer, 1024 beads x 4096 trajectories, PILE"),
    "c2": dict(pes="2dtest", n=256, nintegral=16, nrep=256, thermostat=1, beta=10.0, Noutput=100,
               label="C2: 2D coupled double well TI, 256 beads x 4096 trajectories, Andersen"),
    # C3: the GPU action gradient handed to L-BFGS-B on task 'FG' (instantonmod.f90:748-765); a single ring polymer
    "c3": dict(pes="2dtest", n=1024, beta=30.0, instanton=True,
               label="C3: 2D test PES ring-polymer instanton, 1024 beads, UMforceenergy f/g evaluations (L-BFGS-B on the host)"),
    "c1": dict(pes="1d", n=64, nintegral=16, nrep=16, thermostat=2, beta=10.0, Noutput=100000,
               label="C1: 1D double well TI, 64 beads x 256 trajectories, Langevin"),
}


def wells(pes):
    if pes == "ccpol8sf":
        with open(os.path.join(ROOT, "pimd_tunneling_b200", "data", "ccpol_wells.json")) as f:
            w = json.load(f)
        a = np.asfortranarray(np.array(w["well1"]).T)
        b = np.asfortranarray(np.array(w["well2"]).T)
        return a, b, np.array(w["masses_me"])
    if pes == "2dtest":
        a = np.array([[3.0], [0.0]])  # wells k=6 and k=1 of the C6 ring
        b = np.array([[3.0 * np.cos(np.pi / 3)], [3.0 * np.sin(np.pi / 3)]])
        return np.asfortranarray(a), np.asfortranarray(b), np.array([1.0])
    return np.array([[-1.0]]), np.array([[1.0]]), np.array([1.0])


def ti_path(pes, a, b):
    """(lampath, path, splinepath) as read_path would leave them.  1D/2D: the straight line a->b (the
    natural spline of a line is the line).  Water dimer: a synthetic acceptor-switch
    path (rigid rotation of the acceptor's hydrogens about its bisector; a straight line would drive the
    two hydrogens through each other at lambda = 1/2)."""
    from pimd_tunneling_b200 import path as P

    if pes == "ccpol8sf":
        return P.build_path(P.acceptor_switch_path(a, b, 9))
    pts = np.empty((2,) + a.shape, order="F")
    pts[0], pts[1] = a, b
    return P.build_path(pts)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([s.strip() for s in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 7 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 7 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """one host core: one ring polymer of the workload advanced `nsteps` steps by the oracle"""
    pes, n, beta, thermostat, nsteps, seed, Noutput = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle

    orc = Oracle().select(pes)
    a, b, mass = wells(pes)
    betan = beta / (n + 1)
    orc.nm_setup(n, mass, betan, 1.0, 1.0, 1e-3, False, True)
    orc.set_rng(seed, seed)
    from pimd_tunneling_b200 import path as P

    lam, path, spl = ti_path(pes, a, b)
    xint, dbd = P.endpoints(lam, path, spl, [0.5])
    orc.init_nm(a, xint[..., 0])
    x, p = orc.init_path(0.5, lam, path, spl)
    dbdl = np.asfortranarray(dbd[..., 0])
    t0 = time.perf_counter()
    orc.propagate(thermostat, x, p, dbdl, nsteps, 0, Noutput)
    return time.perf_counter() - t0


def cpu_throughput(cfg, nsteps, cores):
    """bead-steps/s of the oracle with one independent trajectory per core (the reference's MPI model,
    pimd_par.f90:109-110,321-381)"""
    import multiprocessing as mp

    jobs = [(cfg["pes"], cfg["n"], cfg["beta"], cfg["thermostat"], nsteps, 1000 + i, cfg["Noutput"]) for i in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        per = pool.map(_cpu_worker, jobs)
    wall = max(per)
    return cores * cfg["n"] * nsteps / wall, wall, time.perf_counter() - t0


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: one trajectory per core, n beads, a few steps
    per_step = 4 if cfg["pes"] == "ccpol8sf" else 200
    for _ in range(args.warmup if cfg["pes"] != "ccpol8sf" else 1):
        cpu_throughput(cfg, per_step, cores)
    t_tot, units = 0.0, 0
    for _ in range(args.steps):
        v, wall, _ = cpu_throughput(cfg, per_step, cores)
        t_tot += wall
        units += cores * cfg["n"] * per_step
    value = units / t_tot
    sample = "%d cores x 1 trajectory x %d beads x %d oracle step(s) per timed step" % (cores, cfg["n"], per_step)
    line = {
        "impl": "reference", "metric": "ring-polymer bead-steps/sec", "value": value, "unit": "bead-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["label"], "note": "CPU oracle (C++ restatement of the reference algorithm, g++ -O3 -march=native, "
                   "no MKL/ifort: the Fortran reference cannot be compiled in this image)"},
        "cpu_baseline": {"value": value, "unit": "bead-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "bead-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Workload:
    """State of `ntraj` ring polymers of config `cfg` on this rank: pinned host copies (the e2e leg starts from them) and
    device copies (the device-resident leg), end points b(lambda) and dbdl per trajectory, global ids."""

    def __init__(self, pk, cfg, pes, vi, gid, nrep_glob, dev):
        import torch

        from pimd_tunneling_b200 import path as P

        self.cfg, self.vi, self.gid, self.nrep_glob = cfg, vi, gid, nrep_glob
        n, nd, na = cfg["n"], pes.ndim, pes.natom
        ndof = nd * na
        self.n, self.nd, self.na, self.ndof, self.ntraj = n, nd, na, ndof, gid.size
        nintegral = cfg["nintegral"]
        a, b, mass = wells(cfg["pes"])
        self.a = a
        self.xi, self.wts = vi.gauleg(0.0, 1.0, nintegral)
        il = (gid // nrep_glob) % nintegral
        lam, path, spl = ti_path(cfg["pes"], a, b)
        xint, dbdxi = P.endpoints(lam, path, spl, self.xi)       # pimd_par.f90:214-221
        self.bt = np.asfortranarray(xint[:, :, il])              # endpoints(ii,:,:)
        self.dbdl = np.asfortranarray(dbdxi[:, :, il])           # gradpoints(ii,:,:)
        ntraj = self.ntraj
        # init_path in chunks (host staging of the full state is 2 x ntraj*n*ndof*8 bytes)
        self.x_h = torch.empty((ntraj, ndof, n), dtype=torch.float64).pin_memory()
        self.p_h = torch.empty((ntraj, ndof, n), dtype=torch.float64).pin_memory()
        chunk = max(1, min(ntraj, (1 << 27) // (n * ndof)))
        for lo in range(0, ntraj, chunk):
            hi = min(ntraj, lo + chunk)
            xc, pc = vi.init_path(self.xi[il[lo:hi]], lam, path, spl, traj_gid=gid[lo:hi])
            self.x_h[lo:hi] = torch.from_numpy(np.ascontiguousarray(xc.reshape(-1, order="F").reshape(hi - lo, ndof, n)))
            self.p_h[lo:hi] = torch.from_numpy(np.ascontiguousarray(pc.reshape(-1, order="F").reshape(hi - lo, ndof, n)))
        self.x_d, self.p_d = self.x_h.to(dev), self.p_h.to(dev)
        self.a_d = torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).to(dev)
        self.b_d = torch.from_numpy(np.ascontiguousarray(self.bt.reshape(-1, order="F"))).to(dev)
        self.dbdl_d = torch.from_numpy(np.ascontiguousarray(self.dbdl.reshape(-1, order="F"))).to(dev)
        self.gid_d = torch.from_numpy(gid).to(dev)
        self.dH_d = torch.zeros(max(1, ntraj), dtype=torch.float64, device=dev)
        self.sums = None

    def call_dev(self, nsteps, seed):
        """nsteps steps of every local ring polymer (state resident in HBM), then the path's single collective: per-lambda
        estimator sums formed on the device and all-reduced by the library's own NCCL communicator"""
        from pimd_tunneling_b200 import ti

        cfg, vi = self.cfg, self.vi
        vi.seed = seed
        if self.ntraj > 0:
            vi.propagate_dev(cfg["thermostat"], self.ntraj, self.x_d.data_ptr(), self.p_d.data_ptr(), self.a_d.data_ptr(),
                             self.b_d.data_ptr(), self.dbdl_d.data_ptr(), self.gid_d.data_ptr(), self.dH_d.data_ptr(), NMC=nsteps)
        self.sums = ti.reduce_dev(self.ntraj, self.dH_d.data_ptr(), self.gid_d.data_ptr(), self.nrep_glob, cfg["nintegral"], vi.betan)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6500.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def fast_vs_strict(pk, L, check, mode_id, nsample=256):
    """Measured parity of a non-default arithmetic mode against strict on `nsample` thermal dimer geometries:
    max relative energy error and max gradient error over max|grad| (both modes on the GPU, through the C ABI)."""
    rng = np.random.default_rng(3)
    a, _, _ = wells("ccpol8sf")
    x = np.asfortranarray(a[:, :, None] + rng.normal(0.0, 0.05, size=(3, 6, nsample)))
    pes = pk.McmodMass("ccpol8sf").V_init()
    check(L.pimdk_set_mode(0))
    v0, g0 = pes.eval_batch(x)
    check(L.pimdk_set_mode(mode_id))
    v1, g1 = pes.eval_batch(x)
    return {"geometries": nsample, "energy_max_abs_over_max": float(np.max(np.abs(v1 - v0)) / np.max(np.abs(v0))),
            "gradient_max_over_maxgrad": float(np.max(np.abs(g1 - g0).reshape(18, -1).max(0) / np.abs(g0).reshape(18, -1).max(0)))}


MODES = {"strict": 0, "fast": 1, "analytic": 2}


def run_ours(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import pimd_tunneling_b200 as pk
    from pimd_tunneling_b200 import ti
    from pimd_tunneling_b200._lib import check, hptr, lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries the one JSON line; NCCL's version banner (NCCL_DEBUG=VERSION) would go there too
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        # torch.distributed is plumbing here: the barrier of the timing contract and the hand-over of the library
        # communicator's unique id.  The path's collective itself is the library's (pimdk_ti_reduce_dev).
        dist.init_process_group("nccl", device_id=dev)
    pk.init(local_rank)
    L = lib()
    ti.comm_init(rank, world)
    mode_id = MODES[args.mode]
    parity = None
    if mode_id and cfg["pes"] == "ccpol8sf" and rank == 0 and not args.quick and not args.sweep:
        parity = fast_vs_strict(pk, L, check, mode_id)
    check(L.pimdk_set_mode(mode_id if cfg["pes"] == "ccpol8sf" else 0))
    pes = pk.McmodMass(cfg["pes"]).V_init()
    a, b, mass = wells(cfg["pes"])
    n, nd, na = cfg["n"], pes.ndim, pes.natom
    ndof = nd * na
    nintegral = cfg["nintegral"]
    K, W = args.steps, args.warmup
    vi = pk.VerletInt(pes, n, mass, cfg["beta"], dt=1e-3, gamma=1.0, NMC=1, imin=0, Noutput=cfg["Noutput"],
                      seed=0x5EED0000).init_nm()
    strong = args.scaling == "strong"

    def layout(total_or_per_gpu):
        """-> (global ids of this rank, repetitions per lambda of the whole job).  weak: every rank holds `total_or_per_gpu`
        trajectories (the job grows with N); strong: the job is `total_or_per_gpu` trajectories, block-partitioned by
        global id (pimd_par.f90:109-110, 281-295)."""
        if strong:
            nrep_glob = max(1, total_or_per_gpu // nintegral)
            lo, hi = ti.shard(nintegral * nrep_glob, rank, world)
            return np.arange(lo, hi, dtype=np.int64), nrep_glob
        nrep = max(1, total_or_per_gpu // nintegral)
        nt = nintegral * nrep
        return np.arange(nt, dtype=np.int64) + rank * nt, nrep * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(wl, nsteps, seed):
        """device time of one call of nsteps steps on this rank's stream, max over ranks [ms]"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        wl.call_dev(nsteps, seed)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    base_traj = args.ntraj if args.ntraj else nintegral * cfg["nrep"]

    # ---- batch-scaling sweep (C5): one JSON line per batch size, device-resident value only ----
    if args.sweep:
        for tot in [int(t) for t in args.sweep.split(",")]:
            gid, nrep_glob = layout(tot)
            wl = Workload(pk, cfg, pes, vi, gid, nrep_glob, dev)
            wl.call_dev(W, 0x5EED0001)
            barrier()
            check(L.pimdk_profile_reset())
            sampler = ClockSampler(local_rank)
            sampler.start()
            ms = timed(wl, K, 0x5EED0002)
            sampler.stop_flag = True
            ntot = nintegral * nrep_glob
            if rank == 0:
                print(json.dumps({"metric": "ring-polymer bead-steps/sec", "value": ntot * n * K / (ms * 1e-3), "unit": "bead-steps/s",
                                  "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "scaling": args.scaling,
                                  "config": {"workload": cfg["label"], "beads": n, "trajectories_total": ntot,
                                             "trajectories_this_gpu": int(gid.size), "mode": args.mode},
                                  "gpu_launches": int(L.pimdk_launch_count()), "clocks": sampler.summary(), "quick": True}), flush=True)
            del wl
            torch.cuda.empty_cache()
        ti.comm_finalize()
        if world > 1:
            dist.destroy_process_group()
        pk.finalize()
        return

    gid, nrep_glob = layout(base_traj)
    ntraj = int(gid.size)
    ntot = nintegral * nrep_glob                     # trajectories of the whole job
    wl = Workload(pk, cfg, pes, vi, gid, nrep_glob, dev)

    # ---- device-resident timing ("value"): CUDA events on the launching stream, no per-kernel profiling ----
    wl.call_dev(W, 0x5EED0001)                         # W untimed warm-up steps
    barrier()
    check(L.pimdk_profile(0))
    check(L.pimdk_profile_reset())
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(wl, K, 0x5EED0002)                      # exactly K timed steps
    sampler.stop_flag = True
    launches = int(L.pimdk_launch_count())
    units_total = ntot * n * K
    value = units_total / (ms * 1e-3)
    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": "ring-polymer bead-steps/sec", "value": value, "unit": "bead-steps/s", "n_gpus": world,
                              "steps": K, "warmup": W, "ms_per_step": ms / K, "scaling": args.scaling,
                              "config": {"workload": cfg["label"], "beads": n, "trajectories_total": ntot,
                                         "trajectories_this_gpu": ntraj, "mode": args.mode},
                              "gpu_launches": launches, "clocks": sampler.summary(), "quick": True}), flush=True)
        ti.comm_finalize()
        if world > 1:
            dist.destroy_process_group()
        pk.finalize()
        return
    # ---- a second pass of K steps with per-kernel-family CUDA events (roofline of the dominant kernels) ----
    check(L.pimdk_profile(1))
    check(L.pimdk_profile_reset())
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    wl.call_dev(K, 0x5EED0003)
    p1.record()
    torch.cuda.synchronize()
    ms_prof = p0.elapsed_time(p1)
    fam = {}
    for f in ("pes", "gemm", "update", "estimator", "fused"):
        m_, c_ = ctypes.c_double(), ctypes.c_int64()
        check(L.pimdk_profile_get(f.encode(), ctypes.byref(m_), ctypes.byref(c_)))
        fam[f] = {"ms": m_.value, "launches": int(c_.value)}
    check(L.pimdk_profile(0))

    # ---- end-to-end through the host-buffer C ABI ("e2e") ----
    xw = wl.x_h.numpy().reshape(-1).reshape((n, nd, na, ntraj), order="F")
    pw = wl.p_h.numpy().reshape(-1).reshape((n, nd, na, ntraj), order="F")
    dH_h = np.zeros(max(1, ntraj))
    Ke = K

    def e2e_call(nsteps):
        if ntraj > 0:
            check(L.pimdk_propagate(cfg["thermostat"], ntraj, hptr(xw), hptr(pw), hptr(a), hptr(wl.bt), hptr(wl.dbdl), 1e-3, 1.0,
                                    nsteps, 0, cfg["Noutput"], 0, 0x5EED0003, hptr(gid), hptr(dH_h)))
        sums = ti.partial_sums(dH_h[:ntraj], gid, nrep_glob, nintegral, vi.betan)
        return ti.allreduce_sums(sums)             # pimdk_ti_allreduce: the library's communicator

    # one untimed single-step call first: the chunked host-buffer path allocates its device buffers and second stream on first use
    e2e_call(1)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    sums = e2e_call(Ke)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = ntot * n * Ke / t_e2e
    stats = ti.finish(sums, wl.wts, vi.betan)
    state_bytes = 2 * ntraj * n * ndof * 8
    small = ndof * 8 + 2 * ntraj * ndof * 8 + ntraj * 8

    if rank == 0:
        peak = ctypes.c_double()
        check(L.pimdk_fp64_peak(ctypes.byref(peak)))
        peak_note = ("DFMA probe measured in this run by pimdk_fp64_peak (csrc/fp64_peak.cu: 8 independent fma chains per thread, "
                     "148 x 8 x 256 threads, 2 flop per fma, best of 5, CUDA events); MEASURED_PEAKS.json has no FP64 entry; "
                     "theoretical 148 SM x 64 lanes x 2 x sm clock")
        rows_per_gpu = ntraj * ndof
        roof = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak.value, "peak_source": peak_note, "traffic": None,
                "other_kernels_ms": fam, "step_ms_under_profiling": ms_prof / K}
        if cfg["pes"] == "ccpol8sf":
            pes_ms = fam["pes"]["ms"]
            cnt, cnt_src = ccpol_counters(args.mode)
            if args.mode == "analytic":
                # this mode does NOT run the reference's arithmetic, so the reference's operation census does not apply:
                # its work is the FP64 flop its own kernels execute (2*DFMA+DMUL+DADD, ncu source page, profiles/)
                flop_bead = cnt["sass_flop_per_bead"] if cnt else None
                kern = "agrad_{prep,sapt,rigid,back}_kernel (analytic-gradient pipeline; opt-in, not the reference's finite difference)"
                fnote = "executed FP64 flop per bead-gradient of the analytic pipeline (ncu); the reference's 36-energy census does not apply"
            else:
                flop_bead = FLOP_PER_BEAD_GRAD["ccpol8sf"]
                kern = "ccpol_{setup,sites,dipind,sapt,rigid,sweep,combine}_kernel_%s (one PES-gradient pipeline)" % args.mode
                fnote = ("source-level FP64 operation census of the REFERENCE's finite-difference gradient (36 energies; "
                         "oracle/opcount.hpp); exp, division and square root count as one operation each")
            achieved = flop_bead * ntraj * n * K / (pes_ms * 1e-3) / 1e12 if (pes_ms > 0 and flop_bead) else None
            roof.update({"achieved": achieved, "frac": achieved / peak.value if achieved else None,
                         "kernel": kern,
                         "kernel_ms_per_step": pes_ms / K, "kernel_launches": fam["pes"]["launches"],
                         "kernel_share_of_step": pes_ms / ms_prof,
                         "flop_per_bead_gradient": flop_bead,
                         "flop_note": fnote,
                         "algorithmic_bytes": 32 * ndof * ntraj * n / 1e9, "counters_source": cnt_src})
            if cnt:
                roof["traffic"] = cnt["dram_bytes_per_bead"] * ntraj * n / 1e9
                roof["traffic_unit"] = "GB per step (all beads of this GPU): ncu dram__bytes_read+write per bead-gradient x beads"
                if pes_ms > 0 and args.mode != "analytic":
                    roof["achieved_sass"] = cnt["sass_flop_per_bead"] * ntraj * n * K / (pes_ms * 1e-3) / 1e12
                    roof["achieved_sass_note"] = "2*DFMA+DMUL+DADD executed per bead-gradient (ncu source page) / kernel time of this run"
        elif fam["fused"]["launches"] > 0:
            # C1: the whole step loop is ONE kernel; per bead-step it runs two n-point transforms per dof (G = T g, x = T(Q + beadvec)),
            # the surface's gradient, kick + 2 rotations + O-step and its share of the estimator
            flop_bead = 4 * ndof * n + FLOP_PER_BEAD_GRAD[cfg["pes"]] + 30 * ndof
            fms = fam["fused"]["ms"]
            achieved = flop_bead * ntraj * n * K / (fms * 1e-3) / 1e12 if fms > 0 else None
            roof.update({"achieved": achieved, "frac": achieved / peak.value if achieved else None,
                         "kernel": "fused_small_kernel (persistent warp per ring polymer, all steps in one launch)",
                         "kernel_ms_per_step": fms / K, "kernel_launches": fam["fused"]["launches"],
                         "kernel_share_of_step": fms / ms_prof, "flop_per_bead_step": flop_bead,
                         "note": "256 ring polymers = 256 warps on 148 SMs: the configuration is latency/occupancy-bound by size; "
                                 "the fraction is the honest statement of that"})
        else:
            # C2: the two n x n transforms per step dominate: 2 * rows * n^2 flop per launch against the FP64 peak
            gms, gl = fam["gemm"]["ms"], fam["gemm"]["launches"]
            flop = 2.0 * rows_per_gpu * n * n * gl
            achieved = flop / (gms * 1e-3) / 1e12 if gms > 0 else None
            hp, hsrc = hbm_peak()
            ums, ul = fam["update"]["ms"], fam["update"]["launches"]
            roof.update({"achieved": achieved, "frac": achieved / peak.value if achieved else None,
                         "kernel": "nm_gemm_pipe_kernel (FP64 tensor-core transform, cp.async 3-stage)",
                         "kernel_note": "Andersen steps on a model surface: the back-transform's launch also evaluates the bead gradient "
                                        "(24 exponentials per bead for the 2D well) and the forward transform's launch applies the kick, "
                                        "the rotation and the collision clocks, both in the epilogue (no PES launch; the step's other update "
                                        "runs inside the estimator kernel), so the time counted here is not all transform; the plain transform alone runs at "
                                        "22.7 TFLOP/s on this shape (profiles/r2_nm_gemm.md)",
                         "kernel_ms_per_step": gms / K, "kernel_launches": gl, "kernel_share_of_step": gms / ms_prof,
                         "flop_per_launch": 2.0 * rows_per_gpu * n * n,
                         "algorithmic_bytes": 32 * ndof * ntraj * n / 1e9,
                         "update_kernels_hbm": {"ms": ums, "launches": ul, "peak_gbs": hp, "peak_source": hsrc,
                                                "note": "nm_update2 / sample_momenta2 / andersen_clock; 32-48 B per element per launch"}})
        line = {
            "metric": "ring-polymer bead-steps/sec", "value": value, "unit": "bead-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["label"], "pes": cfg["pes"], "beads": n, "trajectories_total": ntot, "trajectories_per_gpu": ntraj,
                       "lambda_points": nintegral, "thermostat": "PILE" if cfg["thermostat"] == 2 else "Andersen",
                       "beta": cfg["beta"], "dt": 1e-3, "mode": args.mode,
                       "l2": "state per step (x,p,P,Q,G: %.3f GB) %s the 126 MB L2" % (5 * state_bytes / 2 / 1e9, "exceeds" if 5 * state_bytes / 2 > 126e6 else "fits; inputs of consecutive steps differ (the state evolves), nothing is cached across steps but the tables"),
                       "parallelism": "independent trajectories sharded by global id; 1 ncclAllReduce of %d doubles inside libpimdk (NCCL %s, %d ranks)" % ((3 * nintegral,) + ti.comm_info()[2:0:-1])},
            "e2e": {"value": e2e_value, "unit": "bead-steps/s", "h2d_bytes_per_step": (state_bytes + small) / Ke,
                    "d2h_bytes_per_step": (state_bytes + ntraj * 8) / Ke, "steps_per_call": Ke,
                    "deltaA": stats["deltaA"], "sigmaA": stats["sigmaA"]},
            "gpu_launches": launches,
            "roofline": roof,
            "clocks": sampler.summary(),
        }
        if parity is not None:
            line["config"]["mode_parity_vs_strict"] = parity
        if cfg["pes"] == "ccpol8sf" and roof.get("traffic"):
            # the same step seen from the memory side (not its bound): DRAM bytes of the PES pipeline per step (ncu, profiles/)
            # plus the streamed state traffic, against the measured copy bandwidth of MEASURED_PEAKS.json
            hp, hsrc = hbm_peak()
            gb = roof["traffic"] + 10 * 8 * ndof * ntraj * n / 1e9
            line["roofline_hbm"] = {"bound": "hbm", "achieved": gb / (ms / K * 1e-3), "peak": hp, "unit": "GB/s",
                                    "frac": gb / (ms / K * 1e-3) / hp, "traffic": gb, "peak_source": hsrc,
                                    "note": "supplementary view: the step is FP64-bound (roofline above); its DRAM traffic uses "
                                            "this fraction of the HBM bandwidth"}
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            per_step = 12 if cfg["pes"] == "ccpol8sf" else 400
            v, wall, tot = cpu_throughput(cfg, per_step, cores)
            line["cpu_baseline"] = {"value": v, "unit": "bead-steps/s", "cores": cores, "kind": "port",
                                    "sample": "%d cores x 1 trajectory x %d beads x %d steps (%.1f s)" % (cores, n, per_step, wall)}
        print(json.dumps(line), flush=True)
    ti.comm_finalize()
    if world > 1:
        dist.destroy_process_group()
    pk.finalize()


def run_instanton(args, cfg, rank):
    """C3: f/g evaluations per second of the ring-polymer action (UMforceenergy) through the host-buffer C ABI —
    what `instanton` calls once per L-BFGS-B iteration — and the same through the CPU oracle; plus one full
    optimisation with scipy's L-BFGS-B (same algorithm and settings as the vendored lbfgsb.f: m = 8, factr = 1e6,
    pgtol = eps2 = 1e-5, maxls = 40) driven by the GPU gradient.  A single polymer does not shard: replicas only."""
    if rank != 0:
        return
    import pimd_tunneling_b200 as pk

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    from scipy.optimize import fmin_l_bfgs_b

    n, beta = cfg["n"], cfg["beta"]
    a, b, mass = wells(cfg["pes"])
    K, W = max(args.steps, 200), max(args.warmup, 20)
    line = {"metric": "instanton action f/g evaluations/sec", "unit": "evaluations/s", "n_gpus": 1, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["label"], "beads": n, "beta": beta, "betan": "beta/n (rpi_ser.f90:45)"}}
    # linear-interpolation start (rpi_ser.f90:157-163)
    x0 = np.empty((n, a.shape[0], a.shape[1]), order="F")
    for i in range(n):
        x0[i] = a + (b - a) * i / (n - 1)
    orc = Oracle().select(cfg["pes"])
    orc.nm_setup(n, mass, beta / n, 1.0, 1.0, 1e-3, False, True)
    if args.impl == "reference":
        for _ in range(W):
            orc.UMforceenergy(x0, a, b)
        t0 = time.perf_counter()
        for _ in range(K):
            orc.UMforceenergy(x0, a, b)
        dt = time.perf_counter() - t0
        line.update({"impl": "reference", "value": K / dt, "ms_per_step": 1e3 * dt / K, "gpu_launches": 0,
                     "cpu_baseline": {"value": K / dt, "unit": "evaluations/s", "cores": 1, "kind": "port",
                                      "sample": "%d evaluations of one 1024-bead polymer, 1 core (a single polymer is serial "
                                                "in the reference too; parallelmod's MPI bead sharding is out of scope)" % K},
                     "e2e": {"value": K / dt, "unit": "evaluations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line), flush=True)
        return
    import torch
    from pimd_tunneling_b200._lib import lib

    pk.init(0)
    pes = pk.McmodMass(cfg["pes"]).V_init()
    v0 = pes.V(a)                 # V0 = V(well1) (rpi_ser.f90:95)
    pes.set_V0(v0)
    orc.set_V0(v0)
    im = pk.InstantonMod(pes, mass, beta, n, fixedends=True, rpi=True)
    for _ in range(W):
        im.UMforceenergy(x0, a, b)
    torch.cuda.synchronize()
    l0 = int(lib().pimdk_launch_count())
    t0 = time.perf_counter()
    for _ in range(K):
        g, f = im.UMforceenergy(x0, a, b)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = int(lib().pimdk_launch_count()) - l0
    nbytes = x0.nbytes
    # the solid-angle loop's shape (rpi_par.f90:209-281): many independent polymers per f/g call
    batch = {}
    for npoly in (8, 64, 512):
        X = np.asfortranarray(np.repeat(x0[..., None], npoly, axis=3))
        B = np.asfortranarray(np.repeat(b[..., None], npoly, axis=2))
        for _ in range(5):
            im.UMforceenergy_batch(X, a, B)
        reps = max(20, K // 4)
        tb = time.perf_counter()
        for _ in range(reps):
            im.UMforceenergy_batch(X, a, B)
        batch[str(npoly)] = reps * npoly / (time.perf_counter() - tb)
    # one complete optimisation driven by the GPU gradient, iterate count and final action against the oracle's
    def fg_gpu(v):
        g_, f_ = im.UMforceenergy(v.reshape(x0.shape, order="F"), a, b)
        return f_, g_.reshape(-1, order="F")

    def fg_cpu(v):
        g_, f_ = orc.UMforceenergy(v.reshape(x0.shape, order="F"), a, b)
        return f_, g_.reshape(-1, order="F")

    t1 = time.perf_counter()
    xg, fgpu, ig = fmin_l_bfgs_b(fg_gpu, x0.reshape(-1, order="F"), m=8, factr=1e6, pgtol=1e-5, maxls=40, maxiter=15000)
    t_opt = time.perf_counter() - t1
    xc, fcpu, ic = fmin_l_bfgs_b(fg_cpu, x0.reshape(-1, order="F"), m=8, factr=1e6, pgtol=1e-5, maxls=40, maxiter=15000)
    t0c = time.perf_counter()
    for _ in range(K):
        orc.UMforceenergy(x0, a, b)
    dtc = time.perf_counter() - t0c
    # close the calculation like `program rpi` does: V0 so that the wells sit at zero, fluctuation factor, splitting
    t2 = time.perf_counter()
    rpi = im.rpi_splitting(xg.reshape(x0.shape, order="F"), a, b)
    rpi["seconds"] = time.perf_counter() - t2
    line.update({"value": K / dt, "ms_per_step": 1e3 * dt / K, "gpu_launches": launches, "rpi": rpi,
                 "batched_evaluations_per_s": batch,
                 "e2e": {"value": K / dt, "unit": "evaluations/s", "h2d_bytes_per_step": nbytes + 2 * a.nbytes + mass.nbytes,
                         "d2h_bytes_per_step": nbytes + 8,
                         "note": "host-buffer C ABI call per evaluation, as L-BFGS-B on the host needs it; value == e2e"},
                 "optimisation": {"iterations": int(ig["nit"]), "funcalls": int(ig["funcalls"]), "seconds": t_opt,
                                  "action_betan_UM": float(fgpu) * beta / n, "oracle_iterations": int(ic["nit"]),
                                  "oracle_funcalls": int(ic["funcalls"]), "rel_diff_action": abs(fgpu - fcpu) / abs(fcpu),
                                  "max_abs_diff_path": float(np.abs(xg - xc).max())},
                 "roofline": {"bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                              "note": "1024 beads x 2 dof: ~1e5 flop and 50 KB per evaluation; the call is bound by launch and "
                                      "PCIe round-trip latency, not by any throughput roofline"},
                 "cpu_baseline": {"value": K / dtc, "unit": "evaluations/s", "cores": 1, "kind": "port",
                                  "sample": "%d evaluations, 1 core" % K}})
    print(json.dumps(line), flush=True)
    pk.finalize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--ntraj", type=int, default=0, help="override trajectories per GPU (testing)")
    ap.add_argument("--mode", default="strict", choices=sorted(MODES),
                    help="CCpol arithmetic: strict = the reference's finite-difference gradient, bit-faithful (the headline); "
                         "fast = the same with FMA contraction; analytic = opt-in analytic gradient (NOT the reference's arithmetic)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: trajectories per GPU fixed; strong: the config's trajectories split over the GPUs")
    ap.add_argument("--sweep", default="", help="comma-separated trajectory counts (per GPU if weak, in total if strong): "
                                                 "one JSON line per count, device-resident value only")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="device-resident value only (no profiling, e2e or CPU legs)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    # stdout carries exactly ONE JSON line.  Libraries write banners to the C-level stdout (NCCL prints its version at
    # any NCCL_DEBUG level from VERSION up, WARN included): point file descriptor 1 at stderr for the whole run and keep
    # the real stdout for Python's own prints.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if cfg.get("instanton"):
        run_instanton(args, cfg, rank)
        return
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=real_stdout))
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
