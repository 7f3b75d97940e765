"""pimd_tunneling_b200 — B200-native (sm_100a CUDA) implementation of the ring-polymer hot path of
christophevaillant/pimd-tunneling behind the reference's own module interfaces:

    mcmod_mass.McmodMass      module mcmod_mass   (V_init, V, Vprime, potforce)
    verletint.VerletInt       module verletint    (init_nm, init_path, propagate_pimd_pile/_nm, gauleg)
    instantonmod.InstantonMod module instantonmod (UM, UMprime, UMforceenergy)
    ti, ti_driver             pimd_par.f90 task layout + statistics; a native TI front end (run_ti)
    path                      splines / reaction coordinate of read_path

The C ABI is include/pimdk.h (libpimdk.so); Fortran binds it through fortran/pimdk_mod.f90.
The directory is spelled with an underscore because `pimd-tunneling_b200` is not importable.
"""
from . import _lib, path, ti, ti_driver  # noqa: F401
from ._lib import PimdkError, finalize, init  # noqa: F401
from .instantonmod import InstantonMod  # noqa: F401
from .mcmod_mass import McmodMass  # noqa: F401
from .verletint import ANDERSEN, PILE, VerletInt  # noqa: F401
