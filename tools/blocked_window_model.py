#!/usr/bin/env python
"""Executable model of the v3 invert kernel's algorithm (suzerain_b200/csrc/invert_blocked.cu):
blocked right-looking banded LU (panel width P = 5 = one collocation point) on a sliding
shared-memory window with a row-slot indirection instead of physical row interchanges,
the right-hand side riding along as an extra row (fused U^T sweep of zgbtrs('T')).

    python tools/blocked_window_model.py      # self-test against SciPy LAPACK
"""
import numpy as np


def cabs1(z):
    return abs(z.real) + abs(z.imag)


def blocked_solve_T(N, KL, KU, entry, b, P=5):
    assert N % P == 0 and (KL + 1) % P == 0
    KV = KL + KU
    RW = KL + P + 1                 # matrix row slots; slot RW = RHS
    CW = KV + P + 1                 # column slots
    W = np.zeros((RW + 1, CW), dtype=complex)
    log_of = np.full(RW + 1, -1)    # logical row held by each slot
    # initial window: logical rows 0..RW-1, columns 0..CW-1
    for i in range(RW):
        log_of[i] = i
        for c in range(CW):
            W[i, c % CW] = entry(i, c) if (i < N and c < N and -KL <= c - i <= KU) else 0.0
    for c in range(CW):
        W[RW, c] = b[c] if c < N else 0.0
    y = np.zeros(N, dtype=complex)
    L = np.zeros((N, KL), dtype=complex)
    ipiv = np.zeros(N, dtype=np.int32)
    ju = 0
    for j in range(0, N, P):
        # ---------- panel factorisation (one warp, registers) ----------
        a = np.array([[W[s, (j + m) % CW] for m in range(P)] for s in range(RW + 1)])
        mylog = log_of.copy()
        pivslot = [-1] * P
        is_piv = np.zeros(RW + 1, dtype=bool)
        for k in range(P):
            col = j + k
            hi = min(col + KL, N - 1)
            best, bl, bs = -1.0, None, None
            for s in range(RW):
                if not is_piv[s] and col <= mylog[s] <= hi:
                    m = cabs1(a[s, k])
                    if m > best or (m == best and mylog[s] < bl):
                        best, bl, bs = m, mylog[s], s
            jp = bl - col
            ipiv[col] = col + jp + 1
            top = [s for s in range(RW) if not is_piv[s] and mylog[s] == col][0]
            mylog[top], mylog[bs] = bl, col          # interchange = relabel
            is_piv[bs] = True
            pivslot[k] = bs
            piv = a[bs].copy()
            if piv[k] == 0:
                return None, ipiv, L, col + 1
            ju = max(ju, min(col + KU + jp, N - 1))
            rinv = 1.0 / piv[k]
            for s in range(RW + 1):
                if is_piv[s]:
                    continue
                l = a[s, k] * rinv
                a[s, k] = l
                # zgbtf2 stores column k's multipliers by the row position held right
                # after step k's interchange (later interchanges do not touch them)
                i = mylog[s] - col
                if s < RW and 1 <= i <= KL and mylog[s] < N:
                    L[col, i - 1] = l
                for m in range(k + 1, P):
                    a[s, m] -= l * piv[m]
        # multipliers / metadata out
        Lp = a.copy()
        for k in range(P):
            Lp[pivslot[k], k:] = 0.0                 # keep only the L11 part of pivot rows
        for m in range(P):
            y[j + m] = a[RW, m]
        # ---------- trailing update: columns j+P .. ju ----------
        for c in range(j + P, ju + 1):
            cs = c % CW
            u = [W[pivslot[m], cs] for m in range(P)]
            for k in range(1, P):
                for m in range(k):
                    u[k] -= Lp[pivslot[k], m] * u[m]
            for s in range(RW + 1):
                if is_piv[s]:
                    continue
                w = W[s, cs]
                for m in range(P):
                    w -= Lp[s, m] * u[m]
                W[s, cs] = w
        # ---------- refresh: rows j+RW..j+RW+P-1 and columns j+CW..j+CW+P-1 enter ----------
        log_of = mylog
        for m in range(P):
            cs = (j + m) % CW            # == (j + CW + m) % CW
            cn = j + CW + m
            for s in range(RW):
                W[s, cs] = 0.0
            W[RW, cs] = b[cn] if cn < N else 0.0
        for k in range(P):
            s, rn = pivslot[k], j + RW + k
            log_of[s] = rn
            for c in range(j + P, j + P + CW):
                W[s, c % CW] = entry(rn, c) if (rn < N and c < N and -KL <= c - rn <= KU) else 0.0
    x = y.copy()
    for j in range(N - 2, -1, -1):
        lm = min(KL, N - 1 - j)
        x[j] -= np.dot(L[j, :lm], x[j + 1:j + 1 + lm])
        l = ipiv[j] - 1
        if l != j:
            x[l], x[j] = x[j], x[l]
    return x, ipiv, L, 0


def _selftest():
    from scipy.linalg import lapack
    rng = np.random.default_rng(7)
    for (N, KL, KU, dom) in [(40, 4, 4, 0.0), (60, 9, 9, 0.0), (120, 14, 14, 2.0), (25, 9, 9, 0.0),
                             (10, 14, 14, 0.0), (5, 4, 4, 0.0), (75, 14, 9, 0.0), (80, 4, 9, 0.0)]:
        A = np.zeros((N, N), dtype=complex)
        for i in range(N):
            for c in range(max(0, i - KL), min(N, i + KU + 1)):
                A[i, c] = rng.standard_normal() + 1j * rng.standard_normal()
            A[i, i] += dom * 4
        b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        ab = np.zeros((2 * KL + KU + 1, N), dtype=complex)
        for i in range(N):
            for c in range(max(0, i - KL), min(N, i + KU + 1)):
                ab[KL + KU + i - c, c] = A[i, c]
        lu, piv, info = lapack.zgbtrf(ab, KL, KU)
        xr, _ = lapack.zgbtrs(lu, KL, KU, b, piv, trans=1)
        x, ipiv, L, info3 = blocked_solve_T(N, KL, KU, lambda i, c: A[i, c], b)
        assert info3 == 0
        assert np.array_equal(ipiv - 1, piv), (N, KL, KU)
        err = np.abs(x - xr).max() / np.abs(xr).max()
        # multipliers agree with LAPACK's factor storage
        kv = KL + KU
        Lref = np.array([[lu[kv + i, jj] if jj + i < N else 0 for i in range(1, KL + 1)] for jj in range(N)])
        lerr = np.abs(L - Lref).max()
        print(f"N={N} KL={KL} KU={KU}: pivots identical ({(piv != np.arange(N)).sum()} non-trivial), "
              f"x relerr {err:.2e}, L abs err {lerr:.2e}")
        assert err < 1e-9 and lerr < 1e-9


if __name__ == "__main__":
    _selftest()
