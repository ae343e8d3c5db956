"""Load a Suzerain restart / fixture file (``fields/*.h5``; written by ``support::save_*`` through ESIO,
``suzerain/support/support.cpp``, ``driver_base.cpp:1734-1795``) into what the implicit operator needs:
grid, scenario, B-spline operators, mean profiles and the wave-space state.  Pure host code on top of
``h5lite`` (no libhdf5 in the image).

Restart fields and the ``bar_*`` samples are B-spline COEFFICIENTS in y; a field is stored as a
``(Fz, Fx, Fy) = (Nz, Nx/2+1, Ny)`` array of ``double[2]`` = complex: wave space in x and z, NOT dealiased,
"in order" in kz (0, 1, .., -1) and only the non-negative kx of the real-to-complex transform
(``support::save_coefficients``, suzerain/support/field.cpp:132-135).  ``Restart.state`` places those modes
into the dealiased wave space the operator walks, as ``support::load_coefficients`` does with
``inorder::wavenumber_translate`` (field.cpp:184-207, suzerain/inorder.c:75-151).
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
    fields: dict                    # name -> complex (Nz, Nx/2+1, Ny) B-spline coefficients, in order in kz
    samples: dict                   # bar_* -> (components, Ny) B-spline coefficients

    def bsplineop(self):
        """The collocation operators rebuilt on the stored breakpoints (support::create_bsplines)."""
        from .api import BsplineOp
        return BsplineOp.from_breakpoints(self.k, self.breakpoints_y)

    def wave_extents(self):
        """(dNz, dNx/2+1): the dealiased wave-space extents of specification_grid / pencil_grid
        (dN = DAF N; real-to-complex in x)."""
        return int(self.Nz * self.DAFz), int(self.Nx * self.DAFx) // 2 + 1

    def state(self):
        """(dNz * (dNx/2+1), 5, Ny) complex: every stored pencil of the DEALIASED wave space in the operator's
        field order (E, mx, my, mz, rho), z slowest then x, as operator_hybrid_isothermal and szb_wavegrid
        walk it; modes the file does not hold (the dealiasing band) are zero."""
        dNz, nx = self.wave_extents()
        dNx = int(self.Nx * self.DAFx)
        out = np.zeros((dNz, nx, 5, self.Ny), dtype=np.complex128)
        for f, name in enumerate(FIELDS):
            v = self.fields[name]
            Fz, Fx, Fy = v.shape
            assert Fy == self.Ny, "restart projection between different B-spline bases is not implemented"
            # X: Fx complex coefficients stand for 2 (Fx - 1) + (Fx & 1) real ones (field.cpp:188-194)
            xs, xd = wavenumber_translate(2 * (Fx - 1) + (Fx & 1), dNx)
            keep = (xs < Fx) & (xd < nx)                       # only kx >= 0 is stored on either side
            xs, xd = xs[keep], xd[keep]
            zs, zd = wavenumber_translate(Fz, dNz)
            out[np.ix_(zd, xd, [f])] = v[np.ix_(zs, xs)][:, :, None, :]
        return np.ascontiguousarray(out.reshape(dNz * nx, 5, self.Ny))

    def mean_profiles(self, bop=None):
        """Collocation-point values of the stored mean samples: D0 applied to their coefficients."""
        D0 = (bop or self.bsplineop()).dense(0)
        return {k: v @ D0.T for k, v in self.samples.items()}


def _wavenumber(N, i):
    return i if i < N // 2 + 1 else i - N                      # suzerain/inorder.h:92-96


def _valid(N, w):
    return -((N - 1) // 2) <= w <= N // 2                      # inorder.h:128-166,258-262 ((1-N)/2 truncates to zero)


def wavenumber_translate(S, T):
    """Index pairs (source, target) of the modes present in both an in-order extent S and an in-order extent T
    (suzerain_inorder_wavenumber_translate, suzerain/inorder.c:75-151, for the whole target range)."""
    src, dst = [], []
    if S <= T:
        for i in range(T):
            w = _wavenumber(T, i)
            if _valid(S, w):
                src.append(w if w >= 0 else S + w); dst.append(i)
    else:
        for i in range(S):
            w = _wavenumber(S, i)
            if _valid(T, w):
                src.append(i); dst.append(w if w >= 0 else T + w)
    return np.array(src, dtype=np.int64), np.array(dst, dtype=np.int64)


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
