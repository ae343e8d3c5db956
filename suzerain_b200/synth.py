"""Synthetic turbulent-channel / boundary-layer inputs for tests and bench.py.

The reference ships no restart file for its headline configuration
(fields/coarse3D.channel.h5 is absent), so inputs follow SURVEY.md section 8d:

* grid: breakpoints = L * htstretch2(htdelta, 1, linspace(0, 1, Ny-k+2))
  (two-sided channel, jobscripts/channel.perfect.in: Ly=2, htdelta=3, k=8) or
  htstretch1 (one-sided plate);
* scenario Re=3000, Ma=1.5, Pr=0.7, gamma=1.4, alpha=0, beta=0.7;
* reference profiles: smooth channel-like means with Reynolds-stress-like
  second moments, turned into the 26 quantities of
  apps/perfect/perfect.cpp:1296-1388 (formulas of suzerain/rholt.hpp:675-709,
  1481-1489);
* state: complex standard normal with amplitude (1 + kx^2 + kz^2)^(-5/6),
  seed 20261017.

Pure numpy; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import numpy as np

SEED = 20261017
SCENARIO = dict(Re=3000.0, Pr=0.7, Ma=1.5, alpha=0.0, gamma=1.4)
BETA_VISC = 0.7
SMR91_ALPHA = (29.0 / 96.0, -3.0 / 40.0, 1.0 / 6.0)
SMR91_BETA = (37.0 / 160.0, 5.0 / 24.0, 1.0 / 6.0)
SMR91_GAMMA = (8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0)
SMR91_ZETA = (0.0, -17.0 / 60.0, -5.0 / 12.0)


def reference_profiles(y, Ly=2.0, scenario=SCENARIO, one_sided=False, seed=SEED):
    """(26, n) reference profiles in lib.REF_NAMES order at collocation points y."""
    y = np.asarray(y, dtype=np.float64)
    g, Ma = scenario["gamma"], scenario["Ma"]
    eta = y / Ly                                   # 0..1
    if one_sided:
        shape = np.tanh(4.0 * eta)                 # boundary-layer-like
        fl = 4 * eta * np.exp(1 - 4 * eta)
    else:
        shape = 1.0 - (2 * eta - 1.0) ** 8         # channel-like, zero at both walls
        fl = np.sin(np.pi * eta) ** 2
    u = 1.0 * shape
    v = 0.02 * fl * (1 - 2 * eta if not one_sided else 1.0)
    w = 0.05 * shape * np.cos(2 * eta)
    T = 1.0 + 0.35 * (1 - shape) * (0.5 if one_sided else 1.0) + 0.2 * shape * (1 - shape)
    rho = 1.0 / T
    # second moments: mean products plus Reynolds-stress-like fluctuations
    uu = u * u + 0.020 * fl
    vv = v * v + 0.004 * fl
    ww = w * w + 0.008 * fl
    uv = u * v - 0.003 * fl * (1 - 2 * eta if not one_sided else 1.0)
    uw = u * w + 0.001 * fl
    vw = v * w + 0.0005 * fl
    u2 = uu + vv + ww
    p = rho * T / g
    m = np.stack([rho * u, rho * v, rho * w])
    e = p / (g - 1) + 0.5 * Ma * Ma * rho * u2
    mu = T ** BETA_VISC
    if one_sided:
        mu = mu.copy()
        mu[-1] = 0.0                               # Redmine #2983: inviscid freestream row
    nu = mu / rho
    e_gradrho = ((g - 2) * e - 2 * p) / (rho * rho) * m
    e_divm = (e + p) / rho
    e_deltarho = mu / (rho * rho) * ((g - 1) * e - 2 * p)
    refs = np.stack([
        u, v, w, u2, uu, uv, uw, vv, vw, ww,
        nu, nu * u, nu * v, nu * w, nu * u2, nu * uu, nu * uv, nu * uw, nu * vv, nu * vw, nu * ww,
        e_gradrho[0], e_gradrho[1], e_gradrho[2], e_divm, e_deltarho])
    assert refs.shape == (26, len(y))
    return np.ascontiguousarray(refs)


def reference_profiles_from_means(rho, u, v, w, T, mu, scenario=SCENARIO):
    """The 26 quantities from pointwise MEAN profiles of a state without fluctuations (the one-dimensional
    restart files of the reference: second moments are products of means), same formulas as above."""
    g, Ma = scenario["gamma"], scenario["Ma"]
    rho, u, v, w, T, mu = (np.asarray(a, dtype=np.float64) for a in (rho, u, v, w, T, mu))
    uu, vv, ww, uv, uw, vw = u * u, v * v, w * w, u * v, u * w, v * w
    u2 = uu + vv + ww
    p = rho * T / g
    m = np.stack([rho * u, rho * v, rho * w])
    e = p / (g - 1) + 0.5 * Ma * Ma * rho * u2
    nu = mu / rho
    e_gradrho = ((g - 2) * e - 2 * p) / (rho * rho) * m
    e_divm = (e + p) / rho
    e_deltarho = mu / (rho * rho) * ((g - 1) * e - 2 * p)
    refs = np.stack([
        u, v, w, u2, uu, uv, uw, vv, vw, ww,
        nu, nu * u, nu * v, nu * w, nu * u2, nu * uu, nu * uv, nu * uw, nu * vv, nu * vw, nu * ww,
        e_gradrho[0], e_gradrho[1], e_gradrho[2], e_divm, e_deltarho])
    return np.ascontiguousarray(refs)


def isothermal_walls(one_sided=False):
    """specification_isothermal-like wall data: (enforce_lower, enforce_upper, lower, upper)
    with lower/upper = (T, u, v, w)."""
    return dict(enforce_lower=True, enforce_upper=not one_sided,
                lower=(1.35, 0.0, 0.0, 0.0), upper=(1.35, 0.0, 0.0, 0.0))


def nrbc_matrices(seed=SEED):
    """Smooth, well-conditioned stand-ins for the Giles matrices
    upper_nrbc_{a,b,c} (5x5, column-major flattening)."""
    rng = np.random.default_rng(seed + 7)
    a = 0.3 * rng.standard_normal((5, 5))
    b = 0.3 * rng.standard_normal((5, 5))
    c = 0.2 * rng.standard_normal((5, 5)) + 0.5 * np.eye(5)
    return (np.asfortranarray(a).reshape(-1, order="F"),
            np.asfortranarray(b).reshape(-1, order="F"),
            np.asfortranarray(c).reshape(-1, order="F"))


def state(km, kn, n, seed=SEED, dtype=np.complex128):
    """(npencil, 5, n) complex state with per-mode amplitude (1+km^2+kn^2)^(-5/6)."""
    km = np.asarray(km, dtype=np.float64)
    kn = np.asarray(kn, dtype=np.float64)
    rng = np.random.default_rng(seed)
    npencil = km.shape[0]
    x = rng.standard_normal((npencil, 5, n)) + 1j * rng.standard_normal((npencil, 5, n))
    amp = (1.0 + km * km + kn * kn) ** (-5.0 / 6.0)
    x *= amp[:, None, None]
    zero = (km == 0) & (kn == 0)
    x[zero] = x[zero].real                         # the mean mode is real
    return x.astype(dtype)


def delta_t(Ly=2.0):
    return 1e-3 * Ly


CONFIGS = {
    # name: (Nx, Ny, Nz, k, htdelta, one_sided)
    "channel_192x96x192": (192, 96, 192, 8, 3.0, False),
    "bl_1024x256x512": (1024, 256, 512, 8, -2.0, True),
    "channel_1536x384x1152": (1536, 384, 1152, 8, 3.0, False),
    "tiny_16x24x16": (16, 24, 16, 6, 2.0, False),
}
LX, LZ = 4 * np.pi, 4 * np.pi / 3
