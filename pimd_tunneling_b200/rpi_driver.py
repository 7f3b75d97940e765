"""Native front end for a ring-polymer-instanton run: what `program rpi` (rpi_ser.f90 / rpi_par.f90) does around the hot
path — namelist RPIDATA with the reference's defaults (rpi_ser.f90:22-43), betan = beta/n (:45), alignment of the wells
(:95-103 like pimd_par), V0 = V(well1), initial guess from the splined path.xyz (read_path) or by linear interpolation
(:157-163), `instanton` (L-BFGS-B on the GPU action gradient), the V0 shift to the lowest bead (:186-192), Q_0 from the
single-well determinant (:214-237), and either the fluctuation factor / kink action / splitting (:350-381) or the
solid-angle loop (:240-340, rpi_par.f90:209-300).  All numerical work is done by libpimdk.so."""
from dataclasses import dataclass, field

import numpy as np

from . import path as P
from .instantonmod import InstantonMod
from .mcmod_mass import McmodMass
from .ti_driver import read_namelist


@dataclass
class RPIData:
    """namelist /RPIDATA/ with the defaults of rpi_ser.f90:28-43"""
    n: int = 100
    beta: float = 30.0
    ndim: int = 3
    natom: int = 1
    xunit: int = 1
    npath: int = 0
    npoints: int = 10
    angular: bool = False
    cutofftheta: float = 2.0 * np.pi
    cutoffphi: float = np.pi
    output_instanton: bool = False
    readpath: bool = True
    alignwell: bool = False
    fixedends: bool = True
    alignpath: bool = True
    checkhess: bool = False
    atom1: int = 1
    atom2: int = 2
    atom3: int = 3
    extra: dict = field(default_factory=dict)


def read_rpidata(text):
    return read_namelist(text, group="RPIDATA", into=RPIData())


def run_rpi(pes_name, rd, well1, well2, mass, path_points=None, path_file=None):
    """One instanton calculation.  Returns a dict: xtilde (the instanton), Vpath along it, lampath, V0, lndetj0 and
    either the splitting quantities of InstantonMod.rpi_splitting or, with rd.angular, the solid-angle table."""
    pes = McmodMass(pes_name).V_init()
    well1 = np.asfortranarray(well1, dtype=np.float64)
    well2 = np.asfortranarray(well2, dtype=np.float64)
    atoms = (rd.atom1, rd.atom2, rd.atom3)
    fixedends = True if rd.angular else rd.fixedends                # rpi_ser.f90:53
    if pes.ndim == 3:
        well1, well2 = P.align_wells(well1, well2, rd.alignwell, atoms)
    pes.set_V0(0.0)
    pes.set_V0(pes.V(well1))                                        # :95
    n = rd.n
    im = InstantonMod(pes, mass, rd.beta, n, fixedends=fixedends, rpi=True)
    if path_file is not None:
        path_points = P.read_xyz_frames(path_file, pes.ndim, pes.natom, rd.xunit)
    if path_points is not None:
        xtilde = P.read_path(path_points, pes.V_batch, n, align=rd.alignpath, atoms=atoms)["xtilde"]
    else:                                                           # "quick and dirty linear interpolation" :157-163
        xtilde = np.empty((n, pes.ndim, pes.natom), order="F")
        for k in range(1, n + 1):
            xtilde[k - 1] = ((n - k) * well1 + (k - 1) * well2) / (n - 1.0)
    xtilde = im.instanton(xtilde, well1, well2) if fixedends else im.instanton(xtilde)
    vpath = pes.V_batch(np.asfortranarray(np.moveaxis(xtilde, 0, 2)))
    if vpath.min() < 0.0:                                           # :186-192: the lowest bead becomes the zero
        i = int(np.argmin(vpath))
        pes.set_V0(0.0)
        pes.set_V0(pes.V(xtilde[i]))
        vpath = pes.V_batch(np.asfortranarray(np.moveaxis(xtilde, 0, 2)))
    lam = np.zeros(n)
    for i in range(1, n):
        lam[i] = lam[i - 1] + np.sqrt(np.sum((xtilde[i] - xtilde[i - 1]) ** 2))
    lam = lam / lam[-1]
    out = {"xtilde": xtilde, "Vpath": vpath, "lampath": lam, "V0": pes.V0, "well1": well1, "well2": well2}
    if rd.angular:
        out["angular"] = im.angular_sweep(xtilde, well1, well2, rd.npoints, rd.cutofftheta, rd.cutoffphi)
    else:
        out.update(im.rpi_splitting(xtilde, well1, well2))
    return out


def crossover(pes, tstate, mass):
    """`program crossover` (crossover.f90): mass-weighted Hessian of the transition state (Vdoubleprime on the GPU),
    its eigenvalues, and the crossover inverse temperature 2 pi / sqrt(-eta2_1) below which the instanton exists.
    (The reference fills hessmat with `do j2 = 1, ndim` where natom is meant (crossover.f90:51), so for natom != ndim
    part of its matrix is never assigned; the full matrix is formed here.)  Returns (beta_c, etasquared)."""
    mass = np.asarray(mass, dtype=np.float64).reshape(pes.natom)
    h = pes.Vdoubleprime(np.asarray(tstate, dtype=np.float64).reshape(pes.ndim, pes.natom))
    nd = pes.ndim * pes.natom
    m = np.empty((nd, nd))
    for i1 in range(pes.ndim):
        for j1 in range(pes.natom):
            for i2 in range(pes.ndim):
                for j2 in range(pes.natom):
                    m[j1 * pes.ndim + i1, j2 * pes.ndim + i2] = h[i1, j1, i2, j2] / np.sqrt(mass[j1] * mass[j2])
    eta = np.linalg.eigvalsh(0.5 * (m + m.T))
    if not eta[0] < 0.0:
        raise ValueError("no negative Hessian eigenvalue: not a transition state")
    return 2.0 * np.pi / np.sqrt(-eta[0]), eta
