"""SciPy restatement of the reference's B-spline collocation operators.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  * suzerain/support/support.cpp:288-311   breakpoints from htstretch1/2
  * suzerain/htstretch.c:40-53,112-125     htstretch1, htstretch2
  * suzerain/bsplineop.c:403-517           bandwidth trimming by exact-zero scan
                                           of the upper-left / lower-right k x k
                                           corners
  * suzerain/bsplineop.c:553-652           D_T[d][offset(ld,kl,ku, j, i)] =
                                           B_j^(d)(xi_i), xi = Greville abscissae
  * suzerain/bsplineop.c:163-187           common ld, max_kl/max_ku views

GSL (gsl_bspline_basis_deriv) is absent from this image; the basis functions
are evaluated with scipy.interpolate.BSpline instead.  Pinned against the
golden collocation matrices of tests/test_bsplineop.cpp (tests/golden).
"""
from __future__ import annotations

import dataclasses

import numpy as np
from scipy.interpolate import BSpline


def htstretch1(delta: float, L: float, x):
    x = np.asarray(x, dtype=np.float64)
    if delta == 0.0:
        return x / L
    return 1 + np.tanh(delta * (x / L - 1)) / np.tanh(delta)


def htstretch2(delta: float, L: float, x):
    x = np.asarray(x, dtype=np.float64)
    if delta == 0.0:
        return x / L
    return 0.5 * (1 + np.tanh(delta * (x / L - 0.5)) / np.tanh(delta / 2))


def breakpoints(ndof: int, k: int, left: float, right: float, htdelta: float):
    """support.cpp:288-300 -- htdelta >= 0 two-sided, < 0 one-sided."""
    b = np.linspace(0.0, 1.0, ndof - k + 2)
    b = htstretch2(+htdelta, 1.0, b) if htdelta >= 0 else htstretch1(-htdelta, 1.0, b)
    return (right - left) * b + left


@dataclasses.dataclass
class BsplineOp:
    """Mirror of suzerain_bsplineop_workspace (bsplineop.h:125-180)."""
    k: int
    n: int
    nderiv: int
    kl: np.ndarray          # per-derivative sub-diagonals of D (not D^T)
    ku: np.ndarray
    max_kl: int
    max_ku: int
    ld: int
    storage: np.ndarray     # (nderiv+1, n, ld) float64: column i of D_T[d] is storage[d, i, :]
    knots: np.ndarray
    greville: np.ndarray

    def D_T_offset(self, d: int) -> int:
        """Element offset of the D_T[d] pointer inside its ld*n block."""
        return int(self.max_ku - self.ku[d])

    def dense(self, d: int) -> np.ndarray:
        """Dense n x n D^(d): [i, j] = B_j^(d)(xi_i)."""
        out = np.zeros((self.n, self.n))
        # view with max bandwidths: D_T[d] - (max_ku - ku[d]) == block start
        for i in range(self.n):          # column i of D^T == row i of D
            for j in range(max(0, i - self.max_ku), min(self.n, i + self.max_kl + 1)):
                out[i, j] = self.storage[d, i, self.max_ku + j - i]
        return out


def _basis_derivs(t, k, x, nderiv):
    """All n basis functions' 0..nderiv derivatives at points x -> (nderiv+1, len(x), n)."""
    n = len(t) - k
    spl = BSpline(t, np.eye(n), k - 1, extrapolate=False)
    out = np.empty((nderiv + 1, len(x), n))
    # Evaluate exactly at the right end point inside the last non-empty span.
    xe = np.array(x, dtype=np.float64)
    for d in range(nderiv + 1):
        if d > k - 1:                       # derivative order beyond the degree: identically zero
            out[d] = 0.0
            continue
        s = spl if d == 0 else spl.derivative(d)
        v = s(xe)
        # extrapolate=False yields nan strictly outside; the right end is included
        out[d] = np.nan_to_num(v, nan=0.0)
    return out


def make_bsplineop(k: int, bpts, nderiv: int | None = None) -> BsplineOp:
    bpts = np.asarray(bpts, dtype=np.float64)
    if nderiv is None:
        nderiv = k - 2                      # support.cpp:308
    t = np.concatenate([np.repeat(bpts[0], k - 1), bpts, np.repeat(bpts[-1], k - 1)])
    n = len(t) - k
    # Greville abscissae: mean of k-1 consecutive interior knots
    xi = np.array([t[i + 1:i + k].mean() for i in range(n)])
    xi[0], xi[-1] = bpts[0], bpts[-1]
    B = _basis_derivs(t, k, xi, nderiv)     # [d, i(point), j(basis)]

    # bandwidths of D^T as the reference counts them (kl/ku refer to D^T's
    # storage: entry D^T[j, i] lives at ku + j - i in column i).
    kl = np.full(nderiv + 1, k - 1)
    ku = np.full(nderiv + 1, k - 1)
    kk = min(k, n)
    for d in range(nderiv + 1):
        Dt = B[d].T                          # D^T[j, i]
        ul = Dt[:kk, :kk]
        lr = Dt[n - kk:, n - kk:]

        def diag_abs_sum(m, off):            # off > 0: super-diagonal of D^T
            return np.abs(np.diagonal(m, off)).sum()
        for off in range(k - 1, 0, -1):      # outermost super-diagonal first
            if diag_abs_sum(ul, off) + diag_abs_sum(lr, off) == 0.0:
                ku[d] -= 1
            else:
                break
        for off in range(k - 1, 0, -1):
            if diag_abs_sum(ul, -off) + diag_abs_sum(lr, -off) == 0.0:
                kl[d] -= 1
            else:
                break
    max_kl, max_ku = int(kl.max()), int(ku.max())
    ld = max_kl + 1 + max_ku
    storage = np.zeros((nderiv + 1, n, ld))
    for d in range(nderiv + 1):
        for i in range(n):
            for j in range(max(0, i - ku[d]), min(n, i + kl[d] + 1)):
                storage[d, i, max_ku + j - i] = B[d, i, j]
        # anything nonzero outside the trimmed band is a construction error
        Dt = B[d].T
        jj, ii = np.nonzero(Dt)
        assert np.all((ii - ku[d] <= jj) & (jj <= ii + kl[d])), "nonzero outside band"
    return BsplineOp(k=k, n=n, nderiv=nderiv, kl=kl.astype(np.int32), ku=ku.astype(np.int32),
                     max_kl=max_kl, max_ku=max_ku, ld=ld, storage=storage,
                     knots=t, greville=xi)
