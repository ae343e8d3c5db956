#!/usr/bin/env python
"""Executable model of the v5 (synchronous, speculative) invert kernel (suzerain_b200/csrc/invert_sync.cu).

Blocked right-looking banded LU of A (order N, KL = KU, panels of P = 5 columns) with the right-hand side carried as one more
row (eliminating b^T with the matrix gives y^T = b^T U^-1, the U^T sweep of zgbtrs('T')), followed by the L^T back
substitution with the interchanges undone in reverse.  Per panel t (j = 5 t) the kernel runs three phases separated by CTA
barriers; inside a phase several ROLES (groups of warps) run concurrently:

    P1   f1     F1(t): the 5 x 5 diagonal block, speculated to need NO interchange; publishes L11 (lp rows 0..4), U11, 1/u_kk, |u_kk|
         tail   the tail rows (positions 5 + NMAIN .. RW: three matrix rows and the right-hand side at order 8) of U(t-1)
         ---- barrier of the assembly warps (the tail still reads the retired pivot rows) ----
         asm    A(t-1): the five rows entering the window overwrite the five retired pivot rows; recycled columns refilled
    P2   f2     F2(t): every other row of the panel from what F1 published; checks izamax's rule on the side
         x      X(t): the pivot rows become rows of U in place (columns j+5 .. ju); originals saved
         (if any check failed: restore, exact unblocked zgbtf2 on the panel with row interchanges, X again)
    P3   u      U(t): rank-5 update of the trailing columns, rows at positions 5 .. 5 + NMAIN - 1 only

The model executes exactly this schedule on a dense copy of the band, records every element each role reads or writes, and
reports an element written by one role and touched by another role of the same phase segment as a race; results are
compared with SciPy's LAPACK (zgbtrf pivots, A^T x = b).

    python tools/sync_window_model.py      # self-test
"""
import numpy as np

P = 5


def cabs1(z):
    return abs(z.real) + abs(z.imag)


class Log:
    def __init__(self):
        self.segment, self.role = 0, None
        self.acc = {}                               # (segment, role) -> (reads, writes)

    def _sets(self):
        return self.acc.setdefault((self.segment, self.role), (set(), set()))

    def read(self, name, idx):
        self._sets()[0].add((name, idx))

    def write(self, name, idx):
        self._sets()[1].add((name, idx))

    def races(self):
        out = []
        segs = {}
        for (seg, role), (r, w) in self.acc.items():
            segs.setdefault(seg, []).append((role, r, w))
        for seg, roles in segs.items():
            for a in range(len(roles)):
                for b in range(len(roles)):
                    if a == b:
                        continue
                    clash = roles[a][2] & (roles[b][1] | roles[b][2])
                    if clash:
                        out.append((seg, roles[a][0], roles[b][0], sorted(clash)[:3]))
        return out


class Mat:
    """The window contents, addressed by GLOBAL (row, column); row N is the right-hand side."""

    def __init__(self, name, a, log):
        self.name, self.a, self.log = name, a, log

    def __getitem__(self, idx):
        self.log.read(self.name, idx)
        return self.a[idx]

    def __setitem__(self, idx, v):
        self.log.write(self.name, idx)
        self.a[idx] = v


def solve_T(A, b, KL, nmain=32, taildef=True, asm_barrier=True):
    """x with A^T x = b and LAPACK's ipiv (0-based), by the v5 schedule.  A: dense N x N with bandwidths KL = KU."""
    N = A.shape[0]
    KU, RW = KL, KL + P + 1
    log = Log()
    # rows that have not entered the window yet must not be touched: keep them aside and let "asm" bring them in
    W = np.zeros((N + 1, N), dtype=complex)
    W[N] = b
    win = Mat("win", W, log)
    lp = Mat("lp", np.zeros((RW + 1, P), dtype=complex), log)        # multipliers of the current panel by row position
    pub = Mat("pub", np.zeros((3, P, P), dtype=complex), log)        # U11, reciprocal pivots, |pivot|
    Lg = np.zeros((N, N), dtype=complex)                             # global multiplier scratch, zgbtf2 (unswapped) order
    ipiv = np.arange(N)
    present = 0                                                      # rows < present are in the window

    def enter(upto, role_logged=True):
        nonlocal present
        for r in range(present, min(upto, N)):
            for c in range(max(0, r - KL), min(N, r + KU + 1)):
                win[r, c] = A[r, c]
        present = max(present, min(upto, N))

    def positions(j):                                                # row positions 5 .. RW of panel j -> global rows
        rows = [j + s for s in range(P, RW) if j + s < N]
        return rows + [N]                                            # + the right-hand side

    log.segment, log.role = -1, "init"
    enter(RW)
    ju = 0
    pending = None                                                   # (j, ju) of the panel whose tail rows are still to be applied
    seg = 0
    for j in range(0, N, P):
        nb = min(P, N - j)
        # ------------------------------------------------ P1
        log.segment = seg; seg += 1
        log.role = "f1"
        a = np.array([[win[j + r, j + c] for c in range(nb)] for r in range(nb)])
        bad = False
        for k in range(nb):
            piv = a[k, k]
            bad |= not (1e-140 < cabs1(piv) < 1e140)
            for r in range(k + 1, nb):
                bad |= cabs1(a[r, k]) > cabs1(piv)
                l = a[r, k] / piv if piv != 0 else 0
                a[r, k] = l
                a[r, k + 1:] -= l * a[k, k + 1:]
        for r in range(nb):
            for k in range(nb):
                if r > k:
                    lp[r, k] = a[r, k]
                else:
                    pub[0, r, k] = a[r, k]
        log.role = "tail"
        if pending is not None and taildef:
            pj, pju = pending
            tails = positions(pj)[nmain:]
            for r in tails:
                pos = (r - pj) if r < N else RW
                for c in range(pj + P, pju + 1):
                    v = win[r, c]
                    for k in range(P):
                        v -= lp[pos, k] * win[pj + k, c]
                    win[r, c] = v
        # ---- barrier of the assembly warps: a new segment for "asm", still concurrent with f1 ----
        if asm_barrier:
            f1_acc = log.acc.pop((log.segment, "f1"), (set(), set()))
            log.segment = seg; seg += 1
            log.acc[(log.segment, "f1")] = f1_acc                    # f1 runs across both halves of phase 1
        log.role = "asm"
        if j > 0:
            enter(j + RW)                                            # rows j + RW - 5 .. j + RW - 1 replace the retired pivot rows:
            for r in range(j - P, j):                                # same slots, so the old rows are overwritten
                for c in range(max(0, r - KL), min(N, r + KU + KL + 1)):
                    win[r, c] = np.nan
        # ------------------------------------------------ P2
        log.segment = seg; seg += 1
        log.role = "f2"
        rows = positions(j)
        for r in rows:
            pos = (r - j) if r < N else RW
            v = np.array([win[r, j + c] for c in range(nb)])
            for k in range(nb):
                cand = r < N and (r - j) <= k + KL
                if cand and cabs1(v[k]) > cabs1(pub[0, k, k]):
                    bad = True
                l = v[k] / pub[0, k, k] if pub[0, k, k] != 0 else 0
                for m in range(k + 1, nb):
                    v[m] -= l * pub[0, k, m]
                lp[pos, k] = l
        log.role = "x"
        juc = min(max(ju, j + nb - 1 + KU), N - 1)
        save = {}
        for c in range(j + nb, juc + 1):
            u = [win[j + k, c] for k in range(nb)]
            save[c] = list(u)
            for k in range(1, nb):
                for i2 in range(k):
                    u[k] -= lp[k, i2] * u[i2]
                win[j + k, c] = u[k]
        if bad:
            # ---------------- exact path: barriers around every step (own segments, single role) ----------------
            log.segment = seg; seg += 1
            log.role = "exact"
            for c, u in save.items():
                for k in range(nb):
                    win[j + k, c] = u[k]
            rows_all = [j + s for s in range(RW) if j + s < N] + [N]
            pan = {r: np.array([win[r, j + c] for c in range(nb)]) for r in rows_all}
            swaps = []
            for k in range(nb):
                cands = [r for r in rows_all if r < N and j + k <= r <= j + k + KL]
                best = max(cands, key=lambda r: (cabs1(pan[r][k]), -r))      # first maximum
                if cabs1(pan[best][k]) == 0:
                    raise ZeroDivisionError("singular")
                ipiv[j + k] = best
                swaps.append(best)
                if best != j + k:
                    pan[best], pan[j + k] = pan[j + k], pan[best]
                    juc = min(max(juc, best + KU), N - 1)
                for r in rows_all:
                    if r == N or r > j + k:
                        l = pan[r][k] / pan[j + k][k]
                        if r < N and r > j + k + KL:
                            assert l == 0
                        pan[r][k] = l
                        pan[r][k + 1:] -= l * pan[j + k][k + 1:]
                        if r < N:
                            Lg[r, j + k] = l                          # to the scratch NOW: zgbtf2's unswapped order
            for k in range(nb):
                for m in range(nb):
                    pub[0, k, m] = pan[j + k][m]
            for r in rows_all:
                pos = (r - j) if r < N else RW
                for k in range(nb):
                    lp[pos, k] = pan[r][k] if (r == N or r > j + k) else 0
            # interchanges on the trailing columns, in order; then X
            for c in range(j + nb, juc + 1):
                for k in range(nb):
                    if swaps[k] != j + k:
                        t0, t1 = win[j + k, c], win[swaps[k], c]
                        win[j + k, c] = t1; win[swaps[k], c] = t0
                u = [win[j + k, c] for k in range(nb)]
                for k in range(1, nb):
                    for i2 in range(k):
                        u[k] -= lp[k, i2] * u[i2]
                    win[j + k, c] = u[k]
        ju = max(ju, juc)
        # multipliers to the scratch (the exact path has stored its own, column by column, before later interchanges
        # moved the rows of the working copy)
        if not bad:
            for k in range(nb):
                for s in range(k + 1, RW):
                    if j + s < N and s <= k + KL:
                        Lg[j + s, j + k] = lp.a[s, k]
        y_panel = [lp.a[RW, k] for k in range(nb)]                  # y = b^T U^-1, entries j .. j+nb-1
        # ------------------------------------------------ P3
        log.segment = seg; seg += 1
        rows = positions(j)
        main = rows[:nmain] if taildef else rows
        for w in range(7):                                           # seven warps, columns dealt cyclically
            log.role = f"u{w}"
            for c in range(j + nb + w, ju + 1, 7):
                for r in main:
                    pos = (r - j) if r < N else RW
                    v = win[r, c]
                    for k in range(nb):
                        v -= lp[pos, k] * win[j + k, c]
                    win[r, c] = v
        pending = (j, ju)
        yield_y = y_panel
        if j == 0:
            y = np.zeros(N, dtype=complex)
        y[j:j + nb] = yield_y
    # L^T back substitution with the interchanges undone in reverse (zgbtrs, TRANS = 'T')
    x = y.copy()
    for jj in range(N - 2, -1, -1):
        lm = min(KL, N - 1 - jj)
        x[jj] -= Lg[jj + 1:jj + 1 + lm, jj] @ x[jj + 1:jj + 1 + lm]
        p = ipiv[jj]
        if p != jj:
            x[jj], x[p] = x[p], x[jj]
    return x, ipiv, log.races()


def _selftest():
    from scipy.linalg import lapack
    rng = np.random.default_rng(7)
    for N, KL, boost in ((60, 14, 0.0), (120, 34, 0.0), (95, 34, 3.0), (60, 14, 5.0), (130, 34, 30.0)):
        A = np.zeros((N, N), dtype=complex)
        for i in range(N):
            for jx in range(max(0, i - KL), min(N, i + KL + 1)):
                A[i, jx] = rng.standard_normal() + 1j * rng.standard_normal()
            A[i, i] += 12.0                                          # diagonally dominant: the speculation holds ...
        for i in rng.choice(N - KL - 1, size=int(boost), replace=False) if boost else []:
            A[i + rng.integers(1, KL), i] += 40.0                    # ... except where a sub-diagonal entry is made the pivot
        b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        x, ipiv, races = solve_T(A, b, KL)
        assert not races, races[:3]
        ab = np.zeros((2 * KL + KL + 1, N), dtype=complex)
        for i in range(N):
            for jx in range(max(0, i - KL), min(N, i + KL + 1)):
                ab[KL + KL + i - jx, jx] = A[i, jx]
        lu, piv, info = lapack.zgbtrf(ab, KL, KL)
        assert info == 0
        assert np.array_equal(piv, ipiv), (N, KL, boost, np.flatnonzero(piv != ipiv)[:5])
        want = np.linalg.solve(A.T, b)
        err = np.abs(x - want).max() / np.abs(want).max()
        assert err < 1e-11, (N, KL, err)
        # the schedule without the deferred tail rows gives the same bits
        x2, ipiv2, races2 = solve_T(A, b, KL, taildef=False)
        assert not races2 and np.array_equal(x, x2) and np.array_equal(ipiv, ipiv2)
        # and the barrier between the tail rows and the assembly is needed: without it the assembly overwrites
        # pivot rows the tail update still reads
        if KL + 1 > 32 and N > 2 * (KL + P + 1):                   # (orders whose trailing update has tail rows)
            _, _, races3 = solve_T(A, b, KL, asm_barrier=False)
            assert any({a, bb} == {"asm", "tail"} for _, a, bb, _ in races3)
    return True


if __name__ == "__main__":
    _selftest()
    print("sync_window_model: pivots and solutions match LAPACK, no races in any phase")
