"""Multi-well post-processing of TI results for the water dimer (restates the arithmetic of the
reference's python_utils/multiwell_waterdimer.py:6-16, 20-84): the ratios q/q0 of the five distinct tunnelling
paths at two inverse temperatures are projected on the irreducible representations of the dimer's G16 permutation-
inversion group, and each level follows from the two-temperature tanh relation.  Host-side, a few dozen flops."""
import numpy as np

HARTREE_TO_CM = 219475.0   # the reference's conversion factor (multiwell_waterdimer.py:80)

# permutation classes in the reference's order: E, (12), (34), (ab)(13)(24), (ab)(14)(23), (ab)(1324), (ab)(1423), (12)(34)
CLASSES = ("E", "A", "D", "AG", "AG", "G", "G", "B")
IRREPS = ("A1+", "E+", "B1+", "A2-", "E-", "B2-")
CHARACTERS = np.array([
    [1, 1, 1, 1, 1, 1, 1, 1],        # A1+
    [1, 1, -1, 0, 0, 0, 0, -1],      # E+
    [1, 1, 1, -1, -1, -1, -1, 1],    # B1+
    [1, -1, -1, -1, -1, 1, 1, 1],    # A2-
    [1, -1, 1, 0, 0, 0, 0, -1],      # E-
    [1, -1, -1, 1, 1, -1, -1, 1],    # B2-
], dtype=np.float64)


def class_weights(A, D, AG, G, B):
    """rho over the eight class representatives from the five path ratios q/q0 (identity = 1)"""
    return np.array([1.0, A, D, AG, AG, G, G, B], dtype=np.float64)


def projection_ratios(rho):
    """I_i = sum_j (1 - c_ij) rho_j / sum_j (1 + c_ij) rho_j for every irreducible representation"""
    rho = np.asarray(rho, dtype=np.float64)
    return ((1.0 - CHARACTERS) @ rho) / ((1.0 + CHARACTERS) @ rho)


def two_temperature_level(Ia, Ib, beta1, beta2):
    """delta = 2 (atanh Ib - atanh Ia) / (beta2 - beta1) and the crossing point betabar (Hartree, a.u.)"""
    y1, y2 = np.arctanh(Ia), np.arctanh(Ib)
    with np.errstate(divide="ignore", invalid="ignore"):
        betabar = (beta1 * y2 - beta2 * y1) / (y2 - y1)
    return 2.0 * (y2 - y1) / (beta2 - beta1), betabar


def waterdimer_levels(rho1, rho2, beta1=12000.0, beta2=20000.0):
    """levels (cm^-1) and betabar per irreducible representation from the class weights at beta1 and beta2"""
    I1, I2 = projection_ratios(rho1), projection_ratios(rho2)
    delta, betabar = two_temperature_level(I1, I2, beta1, beta2)
    return {lab: {"level_cm": HARTREE_TO_CM * d, "betabar": bb, "I1": a, "I2": b}
            for lab, d, bb, a, b in zip(IRREPS, delta, betabar, I1, I2)}
