#!/usr/bin/env python
"""Extracts small fixtures from the reference's restart files (fields/*.h5, written by the real
reference build: GSL B-splines, ESIO/HDF5) with suzerain_b200.h5lite -- run in the build container,
where /root/reference exists; the .npz is committed, the files themselves are not copied.

  restart_fixtures.npz   per file <name>: k, Ny, htdelta, Ly, breakpoints_y, knots,
                         collocation_points_y, integration_weights, Dy0T, Dy1T, Dy2T (the reference's
                         own collocation operators in band storage, suzerain/support/support.cpp
                         save_bsplines); for two channel files also the scenario scalars, the mean
                         profiles bar_{rho,u,T,mu} and the (0,0) mode of the five conserved fields
                         (B-spline coefficients, as restart files store them).

Usage: python tests/golden/make_restart_golden.py [/root/reference]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from suzerain_b200.h5lite import H5File       # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
OPERATORS = ["channel_k06", "channel_k07", "channel_k08", "channel_k09", "coleman3k01.00", "coleman5k30.00",
             "spatiotemporal_1e-1_k06", "legacy_r38808", "spatiotemporal_consistent_1e-1_cevisslam2.602_k08"]
STATES = ["channel_k08", "coleman3k01.00"]

out = {"names": np.array(OPERATORS), "state_names": np.array(STATES)}
for name in OPERATORS:
    f = H5File(os.path.join(REF, "fields", name + ".h5"))
    for key in ("k", "Ny", "htdelta", "Ly"):
        out[f"{name}/{key}"] = f[key].reshape(-1)[0]
    for key in ("breakpoints_y", "knots", "collocation_points_y", "integration_weights", "Dy0T", "Dy1T", "Dy2T"):
        out[f"{name}/{key}"] = f[key]
    a = f.attrs("Dy0T")
    out[f"{name}/kl"], out[f"{name}/ku"] = int(a["kl"][0]), int(a["ku"][0])
for name in STATES:
    f = H5File(os.path.join(REF, "fields", name + ".h5"))
    for key in ("Re", "Ma", "Pr", "gamma", "alpha", "beta", "Lx", "Lz", "t"):
        out[f"{name}/{key}"] = f[key].reshape(-1)[0]
    for key in ("bar_rho", "bar_u", "bar_T", "bar_mu"):
        out[f"{name}/{key}"] = f[key][0]                    # (components, Ny)
    for key in ("rho", "rho_u", "rho_v", "rho_w", "rho_E"):
        v = f[key]                                          # (Nz, Nx, Ny, 2)
        out[f"{name}/{key}"] = (v[..., 0] + 1j * v[..., 1]).reshape(-1)
path = os.path.join(OUT, "restart_fixtures.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes;", len(out), "arrays")
