"""oracle/port.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain-numpy restatement of the reference's implicit-operator hot path.  Nothing
here is shipped or measured: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline legs may import it.  Each function cites the reference file:line
it follows (paths relative to the reference tree).

Pinned against (tests/test_oracle.py):
  * tests/golden/bsmbsm_perm.json     q / qinv tables        (tests/test_bsmbsm.cpp:48-152)
  * tests/golden/bsmbsm_solve.json    30x30 BSMBSM solve, XR (tests/test_bsmbsm.cpp:698-966)
  * oracle/_ref (the reference's own C sources compiled unmodified, LAPACK from
    OpenBLAS) on seeded synthetic pencils: assembly, accumulate, zgbsv and
    zcgbsvx solves, pivots.
Pivot sequences are pinned by no reference test; they are pinned
here against netlib zgbtf2 semantics as implemented by OpenBLAS's LAPACK.
"""
from __future__ import annotations

import numpy as np

REF_NAMES = (
    "ux uy uz u2 uxux uxuy uxuz uyuy uyuz uzuz nu nuux nuuy nuuz nuu2 "
    "nuuxux nuuxuy nuuxuz nuuyuy nuuyuz nuuzuz ex_gradrho ey_gradrho "
    "ez_gradrho e_divm e_deltarho").split()
E, U, V, W, R = range(5)            # ndx::e, mx, my, mz, rho (apps/perfect)
M, D1, D2 = range(3)


# ---------------------------------------------------------------------------
# suzerain/gbmatrix.h:51-68, suzerain/bsmbsm.h:130-186
# ---------------------------------------------------------------------------
def gb_offset(ld, kl, ku, i, j):
    return j * ld + (ku + i - j)


def gb_in_band(kl, ku, i, j):
    return (j - ku <= i) and (i <= j + kl)


def bsmbsm(S, n, kl, ku):
    d = dict(S=S, n=n, kl=kl, ku=ku, ld=kl + 1 + ku, N=S * n,
             KL=S * (kl + 1) - 1, KU=S * (ku + 1) - 1)
    d["LD"] = d["KL"] + 1 + d["KU"]
    return d


def q(S, n, i):
    return (i % S) * n + i // S


def qinv(S, n, i):
    return (i % n) * S + i // n


def aPxpby(trans, S, n, alpha, x, beta, y):
    """suzerain_bsmbsm_zaPxpby (bsmbsm_aPxpby_complex.def:37-110): 'N': y[k] = x[q(k)]."""
    N = S * n
    k = np.arange(N)
    if trans == "N":
        px = np.asarray(x)[q(S, n, k)]
    else:
        px = np.asarray(x)[qinv(S, n, k)]
    return alpha * px + (beta * np.asarray(y) if beta != 0 else 0)


def zpack(A, ihat, jhat, alpha, b, papt, ldpapt):
    """suzerain_bsmbsm_zpack (bsmbsm_pack.def:37-123): scatter the n x n banded block
    b (band storage ld x n, b[ku+i-j + j*ld]) times alpha into PAP^T, zero-filling the
    part of the global band that block (ihat, jhat) owns but b does not cover.
    papt: (N, ldpapt) array with papt[J, KU+I-J] = PAP^T[I, J]."""
    S, n, kl, ku, ld, KU, KL = (A[k] for k in ("S", "n", "kl", "ku", "ld", "KU", "KL"))
    for j in range(n):
        J = qinv(S, n, jhat * n + j)
        for i in range(n):
            I = qinv(S, n, ihat * n + i)
            if not gb_in_band(KL, KU, I, J):
                continue
            inb = gb_in_band(kl, ku, i, j)
            papt[J, KU + I - J] = alpha * b[j, ku + i - j] if (inb and alpha != 0) else 0.0


# ---------------------------------------------------------------------------
# The linearised operator, block by block (suzerain/rholut_imexop.c:121-502;
# the same blocks, transposed, are assembled by rholut_imexop.def:113-498)
# ---------------------------------------------------------------------------
def operator_terms(s, km, kn):
    """List of (row, col, op, ref name | None, coefficient without phi)."""
    g, gm1, gm3 = s["gamma"], s["gamma"] - 1, s["gamma"] - 3
    ap43, ap13 = s["alpha"] + 4.0 / 3.0, s["alpha"] + 1.0 / 3.0
    Ma2, invRe = s["Ma"] * s["Ma"], 1 / s["Re"]
    invMa2, ginvPr, ginvRePr = 1 / Ma2, s["gamma"] / s["Pr"], s["gamma"] / (s["Re"] * s["Pr"])
    ikm, ikn, km2, kn2 = 1j * km, 1j * kn, km * km, kn * kn
    t = []
    add = lambda i, j, op, ref, c: t.append((i, j, op, ref, c))
    # ---- rho_E row (:121-232) ----
    add(E, E, M, "ux", -g * ikm); add(E, E, M, "uz", -g * ikn); add(E, E, M, "nu", -ginvRePr * (km2 + kn2))
    add(E, E, D1, "uy", -g); add(E, E, D2, "nu", ginvRePr)
    add(E, U, M, "nuux", Ma2 * invRe * ((ginvPr - ap43) * km2 + (ginvPr - 1) * kn2))
    add(E, U, M, "nuuz", -Ma2 * invRe * ap13 * km * kn); add(E, U, M, "e_divm", -ikm)
    add(E, U, D1, "nuuy", Ma2 * invRe * ap13 * ikm); add(E, U, D2, "nuux", Ma2 * invRe * (1 - ginvPr))
    add(E, V, M, "nuuy", Ma2 * invRe * (ginvPr - 1) * (km2 + kn2))
    add(E, V, D1, "nuux", Ma2 * invRe * ap13 * ikm); add(E, V, D1, "nuuz", Ma2 * invRe * ap13 * ikn)
    add(E, V, D1, "e_divm", -1.0); add(E, V, D2, "nuuy", Ma2 * invRe * (ap43 - ginvPr))
    add(E, W, M, "nuux", -Ma2 * invRe * ap13 * km * kn)
    add(E, W, M, "nuuz", Ma2 * invRe * ((ginvPr - 1) * km2 + (ginvPr - ap43) * kn2)); add(E, W, M, "e_divm", -ikn)
    add(E, W, D1, "nuuy", Ma2 * invRe * ap13 * ikn); add(E, W, D2, "nuuz", Ma2 * invRe * (1 - ginvPr))
    add(E, R, M, "nuu2", Ma2 * invRe * (km2 + kn2)); add(E, R, M, "nuuxux", Ma2 * invRe * ap13 * km2)
    add(E, R, M, "nuuxuz", Ma2 * invRe * ap13 * 2 * km * kn); add(E, R, M, "nuuzuz", Ma2 * invRe * ap13 * kn2)
    add(E, R, M, "ex_gradrho", -ikm); add(E, R, M, "ez_gradrho", -ikn)
    add(E, R, M, "e_deltarho", -ginvRePr / gm1 * (km2 + kn2))
    add(E, R, D1, "nuuxuy", -Ma2 * invRe * ap13 * 2 * ikm); add(E, R, D1, "nuuyuz", -Ma2 * invRe * ap13 * 2 * ikn)
    add(E, R, D1, "ey_gradrho", -1.0)
    add(E, R, D2, "nuu2", -Ma2 * invRe); add(E, R, D2, "nuuyuy", -Ma2 * invRe * ap13)
    add(E, R, D2, "e_deltarho", ginvRePr / gm1)
    # ---- rho_u row (:234-305) ----
    add(U, E, M, None, -gm1 * invMa2 * ikm)
    add(U, U, M, "ux", gm3 * ikm); add(U, U, M, "uz", -ikn); add(U, U, M, "nu", -invRe * (ap43 * km2 + kn2))
    add(U, U, D1, "uy", -1.0); add(U, U, D2, "nu", invRe)
    add(U, V, M, "uy", gm1 * ikm); add(U, V, D1, "ux", -1.0); add(U, V, D1, "nu", ap13 * invRe * ikm)
    add(U, W, M, "ux", -ikn); add(U, W, M, "uz", gm1 * ikm); add(U, W, M, "nu", -ap13 * invRe * km * kn)
    add(U, R, M, "u2", -0.5 * gm1 * ikm); add(U, R, M, "uxux", ikm); add(U, R, M, "uxuz", ikn)
    add(U, R, M, "nuux", invRe * (ap43 * km2 + kn2)); add(U, R, M, "nuuz", ap13 * invRe * km * kn)
    add(U, R, D1, "uxuy", 1.0); add(U, R, D1, "nuuy", -ap13 * invRe * ikm); add(U, R, D2, "nuux", -invRe)
    # ---- rho_v row (:307-389) ----
    add(V, E, D1, None, -gm1 * invMa2)
    add(V, U, M, "uy", -ikm); add(V, U, D1, "ux", gm1); add(V, U, D1, "nu", ap13 * invRe * ikm)
    add(V, V, M, "ux", -ikm); add(V, V, M, "uz", -ikn); add(V, V, M, "nu", -invRe * (km2 + kn2))
    add(V, V, D1, "uy", gm3); add(V, V, D2, "nu", ap43 * invRe)
    add(V, W, M, "uy", -ikn); add(V, W, D1, "uz", gm1); add(V, W, D1, "nu", ap13 * invRe * ikn)
    add(V, R, M, "uxuy", ikm); add(V, R, M, "uyuz", ikn); add(V, R, M, "nuuy", invRe * (km2 + kn2))
    add(V, R, D1, "u2", -0.5 * gm1); add(V, R, D1, "uyuy", 1.0)
    add(V, R, D1, "nuux", -ap13 * invRe * ikm); add(V, R, D1, "nuuz", -ap13 * invRe * ikn)
    add(V, R, D2, "nuuy", -ap43 * invRe)
    # ---- rho_w row (:391-461) ----
    add(W, E, M, None, -gm1 * invMa2 * ikn)
    add(W, U, M, "ux", gm1 * ikn); add(W, U, M, "uz", -ikm); add(W, U, M, "nu", -ap13 * invRe * km * kn)
    add(W, V, M, "uy", gm1 * ikn); add(W, V, D1, "uz", -1.0); add(W, V, D1, "nu", ap13 * invRe * ikn)
    add(W, W, M, "ux", -ikm); add(W, W, M, "uz", gm3 * ikn); add(W, W, M, "nu", -invRe * (km2 + ap43 * kn2))
    add(W, W, D1, "uy", -1.0); add(W, W, D2, "nu", invRe)
    add(W, R, M, "u2", -0.5 * gm1 * ikn); add(W, R, M, "uxuz", ikm); add(W, R, M, "uzuz", ikn)
    add(W, R, M, "nuuz", invRe * (km2 + ap43 * kn2)); add(W, R, M, "nuux", ap13 * invRe * km * kn)
    add(W, R, D1, "uyuz", 1.0); add(W, R, D1, "nuuy", -ap13 * invRe * ikn); add(W, R, D2, "nuuz", -invRe)
    # ---- rho row (:463-502) ----
    add(R, U, M, None, -ikm); add(R, V, D1, None, -1.0); add(R, W, M, None, -ikn)
    return t


class Problem:
    """Same interface as oracle.ref.Problem, computed with numpy."""

    def __init__(self, op, scenario, refs, bc=None, nrbc=None):
        self.op, self.n = op, op.n
        self.s = {k: float(scenario[k]) for k in ("Re", "Pr", "Ma", "alpha", "gamma")}
        self.refs = {name: np.asarray(refs[i], dtype=np.float64) for i, name in enumerate(REF_NAMES)}
        self.bc = bc
        self.nrbc = None if nrbc is None else tuple(
            None if m is None else np.asarray(m, dtype=np.float64).reshape(5, 5, order="F") for m in nrbc)
        self.A = bsmbsm(5, op.n, op.max_kl, op.max_ku)
        self.D = [op.dense(d) for d in range(3)]          # dense D^(0..2): [i, j] = B_j^(d)(xi_i)

    # ---- (M + phi L) as 25 dense n x n blocks ----
    def blocks(self, phi, km, kn):
        n = self.n
        L = np.zeros((5, 5, n, n), dtype=np.complex128)
        for (i, j, op, ref, c) in operator_terms(self.s, km, kn):
            d = self.refs[ref] if ref is not None else np.ones(n)
            L[i, j] += (c * d)[:, None] * self.D[op]
        blk = phi * L
        for i in range(5):
            blk[i, i] += self.D[M]                          # mass last (rholut_imexop.c:99-103)
        return blk, L

    def accumulate(self, phi, km, kn, x, beta=0.0, y=None, nthreads=1):
        """suzerain_rholut_imexop_accumulate (rholut_imexop.c:43-547) over a batch."""
        x = np.asarray(x, dtype=np.complex128)
        npencil, n = x.shape[0], self.n
        out = np.zeros_like(x) if y is None else np.array(y, dtype=np.complex128, copy=True)
        for p in range(npencil):
            blk, L = self.blocks(phi, km[p], kn[p])
            xin = x[p].reshape(5, n)
            o = (beta * out[p].reshape(5, n)) if beta != 0 else np.zeros((5, n), dtype=np.complex128)
            phiLx = np.einsum("ijab,jb->ia", phi * L, xin)
            o = o + phiLx + np.einsum("ab,ib->ia", self.D[M], xin)
            if self.nrbc is not None:
                # upper-boundary correction (:510-545)
                a, b, c = self.nrbc
                top = xin[:, n - 1]
                t = np.zeros(5, dtype=np.complex128)
                if a is not None:
                    t -= (1j * km[p] * phi) * (a @ top)
                if b is not None:
                    t -= (1j * kn[p] * phi) * (b @ top)
                if c is not None:
                    t -= c @ phiLx[:, n - 1]
                o[:, n - 1] += t
            out[p] = o.reshape(-1)
        return out

    # ---- P (M + phi L)^T P^T in band storage (rholut_imexop.def:41-597) ----
    def dense_papt(self, phi, km, kn, with_bc=True):
        n, N = self.n, self.A["N"]
        blk, _ = self.blocks(phi, km, kn)
        T = np.zeros((N, N), dtype=np.complex128)           # T = PA^TP^T: T[I, J]
        for i in range(5):
            for j in range(5):
                # block (i, j) of A lands transposed: A^T[(j, yj), (i, yi)] = A[(i, yi), (j, yj)]
                T[j::5, i::5] = blk[i, j].T
        if self.nrbc is not None:
            # lower-right 15 x 5 corner: X <- X - X C^T + [0; 0; C^T - i km phi A^T - i kn phi B^T] (:505-595)
            a, b, c = self.nrbc
            I0, J0 = 5 * (n - 3), 5 * (n - 1)
            X = T[I0:I0 + 15, J0:J0 + 5].copy()
            add = np.zeros((15, 5), dtype=np.complex128)
            if c is not None:
                add -= X @ c.T
                add[10:] += c.T
            if a is not None:
                add[10:] -= 1j * km * phi * a.T
            if b is not None:
                add[10:] -= 1j * kn * phi * b.T
            T[I0:I0 + 15, J0:J0 + 5] = X + add
        if with_bc and self.bc is not None:
            self._enforce(T)
        return T

    def _enforce(self, T):
        """IsothermalPATPTEnforcer::op (apps/perfect/operator_hybrid_isothermal.cpp:470-510)."""
        n, A = self.n, self.A
        walls = ([0] if self.bc.get("enforce_lower", 1) else []) + ([1] if self.bc.get("enforce_upper", 1) else [])
        for wall in walls:
            y = 0 if wall == 0 else n - 1
            irho = qinv(5, n, 4 * n + y)
            for eq in range(4):
                ieq = qinv(5, n, eq * n + y)
                factor = self.bc["E_factor"][wall] if eq == 0 else self.bc["vel_factor"][wall][eq - 1]
                s = T[ieq, ieq] if T[ieq, ieq] != 0 else 1.0
                lo, hi = max(0, ieq - A["KU"]), min(A["N"], ieq + A["KL"] + 1)
                T[lo:hi, ieq] = 0
                T[irho, ieq] = -s * factor
                T[ieq, ieq] = s

    def _rhs_bc(self, b):
        """IsothermalPATPTEnforcer::rhs (:516-525)."""
        n = self.n
        walls = ([0] if self.bc.get("enforce_lower", 1) else []) + ([1] if self.bc.get("enforce_upper", 1) else [])
        for wall in walls:
            y = 0 if wall == 0 else n - 1
            for eq in range(4):
                b[qinv(5, n, eq * n + y)] = 0

    def assemble(self, phi, km, kn, packf=False, with_bc=True):
        A = self.A
        T = self.dense_papt(phi, km, kn, with_bc)
        rows = A["LD"] + (A["KL"] if packf else 0)
        off = A["KL"] if packf else 0
        out = np.full((A["N"], rows), np.nan + 1j * np.nan, dtype=np.complex128)
        for J in range(A["N"]):
            for I in range(max(0, J - A["KU"]), min(A["N"], J + A["KL"] + 1)):
                out[J, off + A["KU"] + I - J] = T[I, J]
        return out

    def invert(self, solver, phi, km, kn, state, extra=None, nthreads=1, want_ipiv=False,
               want_iters=False):
        """Loop body of invert_mass_plus_scaled_operator (operator_hybrid_isothermal.cpp:617-686)."""
        A = self.A
        N, KL, KU = A["N"], A["KL"], A["KU"]
        x = np.array(state, dtype=np.complex128, copy=True)
        npencil = x.shape[0]
        ex = None if extra is None else np.array(extra, dtype=np.complex128, copy=True)
        ipiv_all = np.zeros((npencil, N), dtype=np.int32)
        iters = np.zeros(npencil, dtype=np.int32)
        rc, bad = 0, -1
        for p in range(npencil):
            T = self.dense_papt(phi, km[p], kn[p], True)
            ab = dense_to_band(T, KL, KU)
            afb, ipiv, info = zgbtf2(ab, N, KL, KU)
            ipiv_all[p] = ipiv
            if info and not rc:
                rc, bad = info, p
            if info:
                continue
            vecs = [x[p]] + ([ex[p, c] for c in range(ex.shape[1])] if ex is not None else [])
            for vi, v in enumerate(vecs):
                b = aPxpby("N", 5, self.n, 1.0, v, 0.0, None)
                self._rhs_bc(b)
                if solver == "zgbsv":
                    sol = zgbtrs("T", afb, N, KL, KU, ipiv, b)
                else:
                    sol, it = zcgbsvx("T", ab, afb, N, KL, KU, ipiv, b)
                    if vi == 0:
                        iters[p] = it
                v[:] = aPxpby("T", 5, self.n, 1.0, sol, 0.0, None)
        res = {"x": x, "info": rc, "first_bad": bad}
        if ex is not None:
            res["extra"] = ex
        if want_ipiv:
            res["ipiv"] = ipiv_all
        if want_iters:
            res["iters"] = iters
        return res


# ---------------------------------------------------------------------------
# LAPACK restatements (netlib semantics; the reference calls MKL through
# suzerain/blas_et_al/lapack.c:185-197,261-277)
# ---------------------------------------------------------------------------
def dense_to_band(T, KL, KU):
    """LAPACK general band storage with KL extra rows for fill: ab[KL+KU+I-J, J] = T[I, J]."""
    N = T.shape[0]
    ab = np.zeros((2 * KL + KU + 1, N), dtype=np.complex128)
    for J in range(N):
        lo, hi = max(0, J - KU), min(N, J + KL + 1)
        ab[KL + KU + lo - J: KL + KU + hi - J, J] = T[lo:hi, J]
    return ab


def cabs1(z):
    return np.abs(z.real) + np.abs(z.imag)


def zgbtf2(ab, n, kl, ku):
    """Unblocked banded LU with partial pivoting (netlib zgbtf2): pivot = first maximum of
    |re|+|im| (izamax); interchanges applied to columns j..ju; multipliers stored unswapped;
    ipiv 1-based.  Returns (factored copy, ipiv, info)."""
    ab = np.array(ab, dtype=np.complex128, copy=True)
    kv = kl + ku
    ipiv = np.zeros(n, dtype=np.int32)
    info, ju = 0, 0
    for j in range(n):
        km = min(kl, n - 1 - j)
        col = ab[kv:kv + km + 1, j]
        jp = int(np.argmax(cabs1(col)))                    # argmax returns the first maximum
        ipiv[j] = j + jp + 1
        if col[jp] != 0:
            ju = max(ju, min(j + ku + jp, n - 1))
            if jp != 0:
                for c in range(j, ju + 1):                 # zswap along the row, stride ldab-1
                    r0, r1 = kv + j - c, kv + j + jp - c
                    ab[r0, c], ab[r1, c] = ab[r1, c], ab[r0, c]
            if km > 0:
                ab[kv + 1:kv + km + 1, j] *= 1.0 / ab[kv, j]
                for c in range(j + 1, ju + 1):             # zgeru
                    r = kv + j - c
                    ab[r + 1:r + km + 1, c] -= ab[kv + 1:kv + km + 1, j] * ab[r, c]
        elif info == 0:
            info = j + 1
    return ab, ipiv, info


def zgbtrs(trans, afb, n, kl, ku, ipiv, b):
    """netlib zgbtrs for one right hand side."""
    kv = kl + ku
    x = np.array(b, dtype=np.complex128, copy=True)
    if trans == "N":
        for j in range(n - 1):
            lm = min(kl, n - 1 - j)
            l = ipiv[j] - 1
            if l != j:
                x[l], x[j] = x[j], x[l]
            x[j + 1:j + 1 + lm] -= afb[kv + 1:kv + 1 + lm, j] * x[j]
        for j in range(n - 1, -1, -1):
            x[j] /= afb[kv, j]
            lo = max(0, j - kv)
            x[lo:j] -= afb[kv - (j - lo):kv, j] * x[j]
        return x
    # 'T': U^T y = b (forward), then L^T with the interchanges undone (backward)
    for j in range(n):
        lo = max(0, j - kv)
        x[j] = (x[j] - np.dot(afb[kv - (j - lo):kv, j], x[lo:j])) / afb[kv, j]
    for j in range(n - 2, -1, -1):
        lm = min(kl, n - 1 - j)
        x[j] -= np.dot(afb[kv + 1:kv + 1 + lm, j], x[j + 1:j + 1 + lm])
        l = ipiv[j] - 1
        if l != j:
            x[l], x[j] = x[j], x[l]
    return x


def gbmv_T(ab_plain_band, n, kl, ku, x):
    """y = A^T x for the unfactored matrix given in the (2kl+ku+1)-row layout."""
    kv = kl + ku
    y = np.zeros(n, dtype=np.complex128)
    for j in range(n):
        lo, hi = max(0, j - ku), min(n, j + kl + 1)
        y[j] = np.dot(ab_plain_band[kv + lo - j:kv + hi - j, j], x[lo:hi])
    return y


def zcgbsvx(trans, ab, afb, n, kl, ku, ipiv, b, aiter=1, dmax=5, tolsc=0.0):
    """suzerain_lapackext_zcgbsvx with fact='N', siter<0 (dsgbsvx.def:131-318): iterative
    refinement in double precision; tolsc == 0 => absolute tolerance eps."""
    assert trans == "T" and tolsc == 0.0
    eps = np.finfo(np.float64).eps / 2                     # dlamch('E')
    x = np.zeros(n, dtype=np.complex128)
    r = np.array(b, dtype=np.complex128, copy=True)
    res = np.linalg.norm(r)
    lastres = 3.0 * (res + 1.0)
    tol = eps
    diter = -1
    if dmax >= 0 and res > tol:
        while diter < dmax and res > tol:
            diter += 1
            r = zgbtrs(trans, afb, n, kl, ku, ipiv, r)
            x += r
            r = b - gbmv_T(ab, n, kl, ku, x)
            res = np.linalg.norm(r)
            if diter >= aiter and lastres < res * 2.0:
                break
            lastres = res
    else:
        diter = 0
    if np.isnan(res):
        x[:] = np.nan
    return x, diter


# ---------------------------------------------------------------------------
# Reference profiles from the physical-space state: collect_references
# (apps/perfect/perfect.cpp:1266-1400) with the equation-of-state helpers of suzerain/rholut.hpp.
# Pinned against the known answers of the reference's tests/test_rholut.cpp
# (tests/golden/rholut_known_answers.json, tests/test_oracle.py).
# ---------------------------------------------------------------------------
REFERENCE_QUANTITIES = (                      # apps/perfect/references.hpp:83-128, in row order
    "rho p p2 T a u v w u2 uu uv uw vv vw ww nu nu_u nu_v nu_w nu_u2 nu_uu nu_uv nu_uw nu_vv nu_vw nu_ww "
    "ex_gradrho ey_gradrho ez_gradrho e_divm e_deltarho rhou rhov rhow rhoE rhouu rhouv rhouw rhovv rhovw rhoww "
    "rhoEE").split()


def p_T_mu_lambda(alpha, beta, gamma, Ma, rho, m, e):
    """suzerain/rholut.hpp:769-790 (m: (3, ...) array)."""
    rho_inverse = 1 / rho
    m2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2]
    p = (gamma - 1) * (e - Ma * Ma * rho_inverse * m2 / 2)
    T = gamma * p * rho_inverse
    mu = np.power(T, beta)
    lam = (alpha - 2.0 / 3.0) * mu
    return p, T, mu, lam


def explicit_div_e_plus_p_u_refcoeff_div_m(rho, e, p):
    return (e + p) / rho                                          # suzerain/rholt.hpp:675-681


def explicit_div_e_plus_p_u_refcoeff_grad_rho(gamma, rho, m, e, p):
    return (((gamma - 2) * e - 2 * p) / (rho * rho)) * m          # suzerain/rholt.hpp:701-709


def explicit_mu_div_grad_T_refcoeff_div_grad_rho(gamma, mu, rho, e, p):
    return mu / (rho * rho) * ((gamma - 1) * e - 2 * p)           # suzerain/rholt.hpp:1481-1489


def collect_references(alpha, beta, gamma, Ma, sphys, top_is_inviscid=False, with_abs=False):
    """Sums over (z, x) of the 42 reference quantities at every y: the loop body of collect_references
    (apps/perfect/perfect.cpp:1279-1393) before MPI_Allreduce and the chi scaling.  sphys: (5, Ny, Nz, Nx) real,
    fields in ndx order e, mx, my, mz, rho.  top_is_inviscid: mu = lambda = 0 on the last plane (one-sided grids,
    :1287-1290, 1311-1316).  Returns (42, Ny) [and the sums of absolute values, for error scaling]."""
    sphys = np.asarray(sphys, dtype=np.float64)
    e, mx, my, mz, rho = (sphys[i] for i in range(5))
    m = np.stack([mx, my, mz])
    p, T, mu, lam = p_T_mu_lambda(alpha, beta, gamma, Ma, rho, m, e)
    if top_is_inviscid:
        mu = mu.copy(); mu[-1] = 0.0
    u = m / rho
    ux, uy, uz = u
    u2 = ux * ux + uy * uy + uz * uz
    nu = mu / rho
    eg = explicit_div_e_plus_p_u_refcoeff_grad_rho(gamma, rho, m, e, p)
    q = {
        "rho": rho, "p": p, "p2": p * p, "T": T, "a": np.sqrt(T), "u": ux, "v": uy, "w": uz, "u2": u2,
        "uu": ux * ux, "uv": ux * uy, "uw": ux * uz, "vv": uy * uy, "vw": uy * uz, "ww": uz * uz,
        "nu": nu, "nu_u": nu * ux, "nu_v": nu * uy, "nu_w": nu * uz, "nu_u2": nu * u2,
        "nu_uu": nu * ux * ux, "nu_uv": nu * ux * uy, "nu_uw": nu * ux * uz, "nu_vv": nu * uy * uy,
        "nu_vw": nu * uy * uz, "nu_ww": nu * uz * uz,
        "ex_gradrho": eg[0], "ey_gradrho": eg[1], "ez_gradrho": eg[2],
        "e_divm": explicit_div_e_plus_p_u_refcoeff_div_m(rho, e, p),
        "e_deltarho": explicit_mu_div_grad_T_refcoeff_div_grad_rho(gamma, mu, rho, e, p),
        "rhou": mx, "rhov": my, "rhow": mz, "rhoE": e,
        "rhouu": mx * mx / rho, "rhouv": mx * my / rho, "rhouw": mx * mz / rho, "rhovv": my * my / rho,
        "rhovw": my * mz / rho, "rhoww": mz * mz / rho, "rhoEE": e * e / rho,
    }
    # the reference sums with Kahan compensation (apps/perfect/perfect.hpp:78-86): extended precision here
    ld = np.longdouble
    out = np.stack([np.asarray(q[k], dtype=ld).sum(axis=(1, 2)).astype(np.float64) for k in REFERENCE_QUANTITIES])
    if not with_abs:
        return out
    mag = np.stack([np.abs(q[k]).sum(axis=(1, 2)) for k in REFERENCE_QUANTITIES])
    return out, mag
