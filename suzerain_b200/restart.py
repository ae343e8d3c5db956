"""Load a Suzerain restart / fixture file (``fields/*.h5``; written by ``support::save_*`` through ESIO,
``suzerain/support/support.cpp``, ``driver_base.cpp:1734-1795``) into what the implicit operator needs:
grid, scenario, B-spline operators, mean profiles and the wave-space state.  Pure host code on top of
``h5lite`` (no libhdf5 in the image).

Restart fields and the ``bar_*`` samples are B-spline COEFFICIENTS in y; fields are stored as
``(Nz, Nx, Ny)`` arrays of ``double[2]`` = complex, wave space in x and z (``support::save_coefficients``).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .h5lite import H5File

FIELDS = ("rho_E", "rho_u", "rho_v", "rho_w", "rho")          # ndx::{e, mx, my, mz, rho}


@dataclasses.dataclass
class Restart:
    path: str
    Nx: int
    Ny: int
    Nz: int
    k: int
    htdelta: float
    Lx: float
    Ly: float
    Lz: float
    DAFx: float
    DAFz: float
    t: float
    scenario: dict                  # Re, Pr, Ma, alpha, beta, gamma
    breakpoints_y: np.ndarray
    collocation_points_y: np.ndarray
    operators: dict                 # d -> (Ny, ld) band storage of D_T[d] as the reference stored it
    fields: dict                    # name -> complex (Nz, Nx, Ny) B-spline coefficients
    samples: dict                   # bar_* -> (components, Ny) B-spline coefficients

    def bsplineop(self):
        """The collocation operators rebuilt on the stored breakpoints (support::create_bsplines)."""
        from .api import BsplineOp
        return BsplineOp.from_breakpoints(self.k, self.breakpoints_y)

    def state(self):
        """(Nz*Nx, 5, Ny) complex: the stored wave-space pencils in the operator's field order
        (E, mx, my, mz, rho), z slowest then x, as operator_hybrid_isothermal walks them."""
        per = [self.fields[name].reshape(self.Nz * self.Nx, self.Ny) for name in FIELDS]
        return np.ascontiguousarray(np.stack(per, axis=1))

    def mean_profiles(self, bop=None):
        """Collocation-point values of the stored mean samples: D0 applied to their coefficients."""
        D0 = (bop or self.bsplineop()).dense(0)
        return {k: v @ D0.T for k, v in self.samples.items()}


def load(path) -> Restart:
    f = H5File(path)
    scalar = lambda key, default=None: (f[key].reshape(-1)[0] if key in f else default)
    ops = {}
    d = 0
    while f"Dy{d}T" in f:
        ops[d] = f[f"Dy{d}T"]
        d += 1
    fields = {}
    for name in FIELDS:
        if name in f:
            v = f[name]
            fields[name] = v[..., 0] + 1j * v[..., 1]
    samples = {k: f[k][0] for k in f.keys() if k.startswith("bar_") and len(f.shape(k)) == 3}
    return Restart(path=str(path), Nx=int(scalar("Nx")), Ny=int(scalar("Ny")), Nz=int(scalar("Nz")), k=int(scalar("k")),
                   htdelta=float(scalar("htdelta", 0.0)), Lx=float(scalar("Lx")), Ly=float(scalar("Ly")), Lz=float(scalar("Lz")),
                   DAFx=float(scalar("DAFx", 1.5)), DAFz=float(scalar("DAFz", 1.5)), t=float(scalar("t", 0.0)),
                   scenario={k: float(scalar(k)) for k in ("Re", "Pr", "Ma", "alpha", "beta", "gamma") if k in f},
                   breakpoints_y=f["breakpoints_y"] if "breakpoints_y" in f else None,
                   collocation_points_y=f["collocation_points_y"] if "collocation_points_y" in f else None,
                   operators=ops, fields=fields, samples=samples)
