"""Native front end for a thermodynamic-integration run (row N1 of SURVEY §8f): what `program pimd`
(pimd_par.f90) does around the hot path — namelist MCDATA with the reference's defaults (:60-88), wells and
masses files (:119-165), V0 = V(well1) (:166), spline path (read_path), Gauss-Legendre nodes and end points
(:212-221), the (lambda x repetition) task layout (:243-257, 281-295), ONE batched propagate call instead of
the task loop (:321-381), and the statistics (:397-424) with one all-reduce instead of MPI_Gather (:389).
The wells are aligned like pimd_par.f90:159-165 (get_align/align_atoms, path.py) for ndim = 3 and used as given otherwise.  All
numerical work is done by libpimdk.so."""
import re
from dataclasses import dataclass, field

import numpy as np

from . import path as P
from . import ti
from .mcmod_mass import McmodMass
from .verletint import ANDERSEN, PILE, VerletInt


@dataclass
class MCData:
    """namelist /MCDATA/ with the defaults of pimd_par.f90:60-88 (only the members the hot path reads)"""
    n: int = 100
    beta: float = 100.0
    NMC: int = 5000000
    Noutput: int = 100000
    dt: float = 1e-3
    nintegral: int = 5
    nrep: int = 1
    thermostat: int = 1
    ndim: int = 3
    natom: int = 1
    xunit: int = 1
    imin: int = 0
    tau: float = 1.0
    gamma: float = 1.0
    cayley: bool = False
    fixedends: bool = True
    dHdrlimit: float = -1.0
    alignwell: bool = False
    npath: int = 0
    readpath: bool = True
    instapath: bool = False
    centre: bool = False
    readhess: bool = False
    atom1: int = 1
    atom2: int = 2
    atom3: int = 3
    seed: int = 0
    extra: dict = field(default_factory=dict)


def read_namelist(text, group="MCDATA", into=None):
    """&MCDATA key=value, ... /   (Fortran logicals .true./.false., d-exponents); `group`/`into` select another
    namelist group and its dataclass (rpi_driver reads &RPIDATA with this)"""
    body = re.search(r"&\s*" + group + r"(.*?)(/|&end)", text, flags=re.S | re.I)
    if not body:
        raise ValueError("no &%s namelist found" % group)
    mc = MCData() if into is None else into
    for key, val in re.findall(r"(\w+)\s*=\s*([^,\n/]+)", body.group(1)):
        v = val.strip().strip("'\"")
        lk = {f.lower(): f for f in mc.__dataclass_fields__}
        if key.lower() not in lk:
            mc.extra[key] = v
            continue
        name = lk[key.lower()]
        cur = getattr(mc, name)
        if isinstance(cur, bool):
            setattr(mc, name, v.lower().startswith((".t", "t")))
        elif isinstance(cur, int):
            setattr(mc, name, int(float(v.lower().replace("d", "e"))))
        else:
            setattr(mc, name, float(v.lower().replace("d", "e")))
    return mc


def read_wells(path1, path2, masses_path, ndim, natom, xunit=1):
    """well1.dat / well2.dat: natom lines of ndim numbers; masses.dat: label mass (pimd_par.f90:123-165)"""
    w1 = np.loadtxt(path1, ndmin=2)[:natom, :ndim].T.copy()
    w2 = np.loadtxt(path2, ndmin=2)[:natom, :ndim].T.copy()
    if xunit == 2:
        w1, w2 = w1 / 0.529177, w2 / 0.529177
    labels, mass = [], []
    with open(masses_path) as f:
        for line in f:
            t = line.split()
            if len(t) >= 2:
                labels.append(t[0])
                mass.append(float(t[1].lower().replace("d", "e")))
    return np.asfortranarray(w1), np.asfortranarray(w2), np.array(mass[:natom]), labels[:natom]


def run_ti(pes_name, mc, well1, well2, mass, path_points=None, rank=0, world=1, max_traj_per_call=None, path_file=None):
    """One TI run.  Returns the statistics of pimd_par.f90:397-424 plus the per-trajectory integrands of this
    rank.  With world > 1 (torch.distributed initialised) the estimator sums are all-reduced.
    The path: `path_file` (the reference's path.xyz, read like read_path with mc.xunit, mc.instapath, mc.centre) or
    `path_points` (frames already in bohr) or, with neither, the straight line between the wells."""
    pes = McmodMass(pes_name).V_init()
    well1 = np.asfortranarray(well1, dtype=np.float64)
    well2 = np.asfortranarray(well2, dtype=np.float64)
    atoms = (mc.atom1, mc.atom2, mc.atom3)
    if pes.ndim == 3:   # pimd_par.f90:159-165 (the reference STOPs for ndim != 3; there the wells are used as given)
        well1, well2 = P.align_wells(well1, well2, mc.alignwell, atoms)
    pes.set_V0(0.0)
    pes.set_V0(pes.V(well1))                                   # V0 = V(well1), pimd_par.f90:166
    if path_file is not None:
        path_points = P.read_xyz_frames(path_file, pes.ndim, pes.natom, mc.xunit)
    if path_points is None:
        path_points = np.stack([well1, well2], axis=0)
        lam, path, spl = P.build_path(path_points)
    else:
        refine = None
        if mc.instapath:   # read_path :953-990: the instanton found from the splined guess becomes the path
            from .instantonmod import InstantonMod

            im = InstantonMod(pes, mass, mc.beta, mc.n, fixedends=mc.fixedends, rpi=False)
            refine = (lambda xt: im.instanton(xt, well1, well2)) if mc.fixedends else (lambda xt: im.instanton(xt))
        rp = P.read_path(path_points, pes.V_batch, mc.n, align=True, instanton=refine, well1=well1, well2=well2,
                         fixedends=mc.fixedends, centre=mc.centre, atoms=atoms)
        lam, path, spl = rp["lampath"], rp["path"], rp["splinepath"]
    vi = VerletInt(pes, mc.n, mass, mc.beta, tau=mc.tau, gamma=mc.gamma, dt=mc.dt, NMC=mc.NMC, imin=mc.imin,
                   Noutput=mc.Noutput, cayley=mc.cayley, seed=mc.seed).init_nm()
    xi, weights = vi.gauleg(0.0, 1.0, mc.nintegral)
    xint, dbdxi = P.endpoints(lam, path, spl, xi)
    ids = ti.global_ids(mc.nintegral, mc.nrep)
    lo, hi = ti.shard(ids.size, rank, world)
    gid = ids[lo:hi]
    il = gid // mc.nrep
    startpoint = np.asfortranarray(path[0])                    # startpoint(:,:) = path(1,:,:), :251
    dH = np.empty(gid.size)
    step = max_traj_per_call or gid.size or 1
    for s in range(0, gid.size, step):
        sl = slice(s, min(gid.size, s + step))
        x, p = vi.init_path(xi[il[sl]], lam, path, spl, traj_gid=gid[sl], readhess=mc.readhess)
        b = np.asfortranarray(xint[:, :, il[sl]])
        dbdl = np.asfortranarray(dbdxi[:, :, il[sl]])
        fn = vi.propagate_pimd_pile if mc.thermostat == PILE else vi.propagate_pimd_nm
        # pimd_par.f90:323-326: a task whose end point is all zeros (the padding of the reference's block layout) is not
        # propagated and contributes integrand 0
        live = ~np.all(np.abs(b) < 1e-10, axis=(0, 1))
        out_dH = np.zeros(live.size)
        if live.any():
            if mc.thermostat == PILE:   # dHdrlimit: verletmodule.f90:404-409 (propagate_pimd_nm ignores the limit)
                vi.set_dhdrlimit(mc.dHdrlimit, xi[il[sl]][live], lam, path, spl)
            _, _, out_dH[live] = fn(np.asfortranarray(x[..., live]), np.asfortranarray(p[..., live]), startpoint,
                                    np.asfortranarray(b[..., live]), np.asfortranarray(dbdl[..., live]), traj_gid=gid[sl][live])
        dH[sl] = out_dH
    sums = ti.allreduce_sums(ti.partial_sums(dH, gid, mc.nrep, mc.nintegral, vi.betan))
    out = ti.finish(sums, weights, vi.betan)
    out.update({"xi": xi, "weights": weights, "integrand": dH / vi.betan ** 2, "traj_gid": gid, "betan": vi.betan,
                "V0": pes.V0})
    return out
