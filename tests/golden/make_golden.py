#!/usr/bin/env python3
"""Regenerate tests/golden/*.json from the CPU oracle (which is itself pinned to the reference's only
golden vectors, the 20 known-answer energies of main_CCpol-8sf.f:180-189 — see tests/test_oracle.py).
The reference is Fortran and cannot run here, so these fixtures are oracle outputs, not reference
outputs; they freeze the oracle's bits so that any later change to oracle OR kernels is caught."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Oracle, thermal_dimer_geometries  # noqa: E402


def hexlist(a):
    return [float(v).hex() for v in np.asarray(a).reshape(-1, order="F")]


def main():
    orc = Oracle()
    out = {}
    orc.select("ccpol8sf")
    x = thermal_dimer_geometries(16, seed=2024)
    v, g, xd = orc.pes_eval(x)
    out["ccpol8sf"] = {"x": hexlist(x), "v": hexlist(v), "grad": hexlist(g), "x_after_vprime": hexlist(xd), "nbatch": 16}
    rng = np.random.default_rng(7)
    orc.select("2dtest")
    x2 = np.asfortranarray(rng.normal(0, 2.5, size=(2, 1, 32)))
    v, g, _ = orc.pes_eval(x2)
    out["2dtest"] = {"x": hexlist(x2), "v": hexlist(v), "grad": hexlist(g), "nbatch": 32}
    orc.select("1d")
    x1 = np.asfortranarray(rng.normal(0, 1.5, size=(1, 1, 32)))
    v, g, _ = orc.pes_eval(x1)
    out["1d"] = {"x": hexlist(x1), "v": hexlist(v), "grad": hexlist(g), "nbatch": 32}
    # RNG contract: first normals of a few (stream, step, gid) triples
    out["normals"] = [{"seed": s, "stream": st, "step": sp, "gid": gd, "idx": i,
                       "z": float(orc.L.orc_normal(s, st, sp, gd, i)).hex()}
                      for (s, st, sp, gd) in ((0, 0, 0, 0), (1234, 1, 17, 5), (2 ** 40 + 3, 2, 2 ** 33, 4000000000))
                      for i in (0, 1, 2, 1001)]
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", os.path.join(HERE, "oracle_vectors.json"))


if __name__ == "__main__":
    main()
