#!/usr/bin/env python
"""Executable model of the register-window banded LU used by the v2 invert kernel
(suzerain_b200/csrc/invert_window.cuh).  Pure numpy, thread-granular: it mimics
the CUDA kernel's data ownership (cyclic row/column slots spread over TR x TC
threads), its phase structure and its index algebra, so that the algorithm can
be checked against LAPACK zgbtrf/zgbtrs('T') without a GPU.

    python tools/window_lu_model.py            # self-test against SciPy LAPACK
"""
from __future__ import annotations

import numpy as np


def recip(z):
    return 1.0 / z


def cabs1(z):
    return abs(z.real) + abs(z.imag)


def window_solve_T(N, KL, KU, entry, b, TR=6, TC=24):
    """Solves A^T x = b where A = P L U is the zgbtf2 factorisation of the band
    matrix A (entry(I, J) -> A[I, J]); returns x, ipiv (1-based), L multipliers,
    info.  Mirrors the kernel: forward sweep = LU with the right-hand side riding
    along as an extra row (y^T = b^T U^-1), backward sweep = L^T with the row
    interchanges undone in reverse."""
    R, KV = KL + 1, KL + KU
    C = KV + 1
    RA = -(-(R + 1) // TR)          # row slots per thread (slot R = RHS row)
    CB = -(-C // TC)
    nthreads = TR * TC
    W = np.zeros((nthreads, RA, CB), dtype=complex)
    rs = lambda tid, a: (tid % TR) + a * TR
    cs = lambda tid, b_: (tid // TR) + b_ * TC
    # ownership maps
    own = {}
    for tid in range(nthreads):
        for a in range(RA):
            for b_ in range(CB):
                if rs(tid, a) <= R and cs(tid, b_) < C:
                    own[(rs(tid, a), cs(tid, b_))] = (tid, a, b_)

    def get(r, c):
        t, a, b_ = own[(r, c)]
        return W[t, a, b_]

    def put(r, c, v):
        t, a, b_ = own[(r, c)]
        W[t, a, b_] = v

    # initial window: rows 0..KL, columns 0..KV; RHS row t_c = b_c
    for i in range(min(R, N)):
        for c in range(C):
            put(i % R, c % C, entry(i, c) if (c < N and -KL <= c - i <= KU) else 0.0)
    for c in range(C):
        put(R, c % C, b[c] if c < N else 0.0)

    y = np.zeros(N, dtype=complex)
    L = np.zeros((N, KL), dtype=complex)
    ipiv = np.zeros(N, dtype=np.int32)
    info = 0
    for j in range(N):
        jr, jc = j % R, j % C
        km = min(KL, N - 1 - j)
        # A/B: publish column j and row j
        s_col = np.array([get(r, jc) for r in range(R + 1)])
        s_top = np.array([get(jr, c) for c in range(C)])
        # C: pivot search
        best, jp = -1.0, 0
        for i in range(km + 1):
            m = cabs1(s_col[(jr + i) % R])
            if m > best:
                best, jp = m, i
        ipiv[j] = j + jp + 1
        piv = s_col[(jr + jp) % R]
        if piv == 0:
            info = j + 1
            break
        rp = (jr + jp) % R
        # D: publish pivot row; old top row takes its slot
        s_piv = np.array([get(rp, c) for c in range(C)])
        if jp != 0:
            for c in range(C):
                put(rp, c, s_top[c])
        rinv = recip(piv)
        # F: multipliers + rank-1 update (all threads, own elements)
        for tid in range(nthreads):
            for a in range(RA):
                r = rs(tid, a)
                if r > R:
                    continue
                if r == R:
                    l = s_col[R] * rinv
                else:
                    rel = (r - jr) % R
                    if not (1 <= rel <= km):
                        continue
                    val = s_col[jr] if rel == jp else s_col[r]
                    l = val * rinv
                for b_ in range(CB):
                    c = cs(tid, b_)
                    if c >= C:
                        continue
                    relc = (c - jc) % C
                    if 1 <= relc <= KV:
                        W[tid, a, b_] -= l * s_piv[c]
        for rel in range(1, km + 1):
            r = (jr + rel) % R
            val = s_col[jr] if rel == jp else s_col[r]
            L[j, rel - 1] = val * rinv
        y[j] = s_col[R] * rinv
        # G: retire column j / row j; enter column j+KV+1 and row j+KL+1
        cn, rn = j + KV + 1, j + KL + 1
        for r in range(R):
            put(r, jc, 0.0)
        put(R, jc, b[cn] if cn < N else 0.0)
        for c in range(rn - KL, rn + KU + 1):        # == j+1 .. j+KV+1
            put(jr, c % C, entry(rn, c) if (rn < N and 0 <= c < N) else 0.0)
    if info:
        return None, ipiv, L, info
    # backward: L^T with interchanges undone
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
    rng = np.random.default_rng(5)
    for (N, KL, KU, TR, TC, dom) in [(40, 4, 4, 3, 5, 0.0), (57, 5, 3, 4, 3, 0.0), (120, 14, 14, 6, 8, 1.0),
                                    (23, 9, 9, 5, 7, 0.0), (8, 9, 9, 5, 7, 0.0), (1, 2, 2, 2, 2, 0.0)]:
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
        xr, info2 = lapack.zgbtrs(lu, KL, KU, b, piv, trans=1)
        x, ipiv, L, info3 = window_solve_T(N, KL, KU, lambda i, c: A[i, c], b, TR, TC)
        assert info3 == 0
        assert np.array_equal(ipiv - 1, piv), (N, KL, KU)
        err = np.abs(x - xr).max() / np.abs(xr).max()
        res = np.abs(A.T @ x - b).max()
        print(f"N={N} KL={KL} KU={KU} TRxTC={TR}x{TC}: pivots identical, relerr {err:.2e}, resid {res:.2e},"
              f" nontrivial pivots {(piv != np.arange(N)).sum()}")
        assert err < 1e-10


if __name__ == "__main__":
    _selftest()
