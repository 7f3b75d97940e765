"""Exact value of what a thermodynamic-integration run estimates, by transfer matrices on a grid (test infrastructure).

`program pimd` (pimd_par.f90:397-424) prints q/q0 = exp(-betan * DeltaA), DeltaA = sum_i w_i <dH/dxi>_i / betan^2 with
the estimator of verletmodule.f90:236-244, 397-410.  With H the open-chain ring-polymer potential of
instantonmod.f90:17-46 (fixed ends a and b(xi)),

    q(xi) = int dx_1..dx_n exp(-betan [ sum_i V(x_i) + m/(2 betan^2) ( |x_1-a|^2 + sum_i |x_{i+1}-x_i|^2 + |x_n-b(xi)|^2 ) ])

so q(1)/q(0) is a ratio of n-bead discretised density-matrix elements.  (The estimator omits the x-independent part
m b.db/dxi / betan^2 of dH/dxi; its xi-integral m (|b(1)|^2 - |b(0)|^2) / (2 betan^2) vanishes for wells at equal
distance from the origin, which holds for both model surfaces.)  The same integral is evaluated here
deterministically: v <- F(a, .); n times { v <- v * exp(-betan V); v <- F v } with the free-particle kernel
F(x, x') = exp(-m |x-x'|^2 / (2 betan)) on a uniform grid (trapezoid rule; the integrand is a product of Gaussians
times a smooth function, so the rule converges geometrically in sigma/h).  Nothing here uses the oracle or the
library: it pins BOTH against the mathematics of the method — integrator, thermostat, estimator and statistics."""
import numpy as np


def v_1d(x, vh=1.0, x0=1.0):
    """mcmod_1d.f90:22-30"""
    return vh * ((x / x0) ** 2 - 1.0) ** 2


def v_2d(x, y, a0=2.0, b0=0.2, rho0=3.0, nwell=6):
    """mcmod_2dtest.f90:27-40 (without V0: a constant cancels in the ratio)"""
    v = np.zeros(np.broadcast(x, y).shape)
    for k in range(1, nwell + 1):
        xk, yk = rho0 * np.cos(2 * np.pi * k / nwell), rho0 * np.sin(2 * np.pi * k / nwell)
        u = (x - xk) ** 2 + (y - yk) ** 2
        v = v - 0.5 * (np.exp(-a0 * u) + np.exp(-b0 * u))
    return v


def log_ratio_1d(a, b, n, beta, mass=1.0, lo=-4.0, hi=4.0, h=0.01):
    """ln [ q(xi=1) / q(xi=0) ] for the 1D double well, ends a -> b against a -> a"""
    betan = beta / (n + 1)
    x = np.arange(lo, hi + h / 2, h)
    F = np.exp(-mass * (x[:, None] - x[None, :]) ** 2 / (2 * betan)) * h
    w = np.exp(-betan * v_1d(x))
    v = np.exp(-mass * (x - a) ** 2 / (2 * betan))
    for i in range(n):
        v = v * w
        if i < n - 1:
            v = F @ v
    end = lambda e: np.sum(v * np.exp(-mass * (x - e) ** 2 / (2 * betan))) * h
    return float(np.log(end(b) / end(a)))


def log_ratio_2d(a, b, n, beta, mass=1.0, lim=7.5, h=0.1):
    """ln [ q(1) / q(0) ] for the 2D test surface; a, b are (x, y) pairs"""
    betan = beta / (n + 1)
    g = np.arange(-lim, lim + h / 2, h)
    F = np.exp(-mass * (g[:, None] - g[None, :]) ** 2 / (2 * betan)) * h
    X, Y = np.meshgrid(g, g, indexing="ij")
    w = np.exp(-betan * v_2d(X, Y))
    v = np.exp(-mass * ((X - a[0]) ** 2 + (Y - a[1]) ** 2) / (2 * betan))
    for i in range(n):
        v = v * w
        if i < n - 1:
            v = F @ v @ F.T
    end = lambda e: np.sum(v * np.exp(-mass * ((X - e[0]) ** 2 + (Y - e[1]) ** 2) / (2 * betan))) * h * h
    return float(np.log(end(b) / end(a)))
