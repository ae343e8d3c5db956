#!/usr/bin/env python
"""Executable model of the v4 (pipelined) invert kernel (suzerain_b200/csrc/invert_pipe.cu).

Same blocked right-looking banded LU on a sliding, slot-indirected shared-memory window as
tools/blocked_window_model.py, but scheduled as a two-stage pipeline per panel t (j = 5 t):

    all warps    : X(t-1)  the pivot rows of panel t-1 become rows of U in place (columns j .. ju)
                   U0(t-1) rank-5 update of block t (columns j .. j+4) of the window
    panel warp   : load block t of every row (entering rows from the stage)
                   F(t)    factor panel t in registers, publish multipliers / pivot slots
    update warps : U(t-1)  rank-5 update of columns j+5 .. ju(t-1) in the window
                   R(t-1)  rows entering after panel t-1 replace the retired pivot rows,
                           recycled column slots zeroed
                   A(t)    assemble the rows entering after panel t into the stage
    one CTA barrier per panel

The model executes the two sides of an iteration in both orders and records every shared
array element each side reads or writes; any element written by one side and touched by
the other within the same iteration is reported as a race.

split_u0=True is the schedule the kernel runs when built with `make SPLITU0=1` (DESIGN 7.1a'; correct but
measured slower, so off by default): the
update of block t+1 against pivots 0..3 of panel t ("u0a") moves behind the update warps' own work
of the iteration, concurrent with the panel warps' last column step (multipliers and pivot slots are
published column by column), so that the all-warps stage before F(t+1) shrinks to a rank-one update of
block t+1 with the last pivot plus the in-place U rows of the columns past it -- which touch disjoint
columns and need no barrier between them.  The model checks it for races and against LAPACK.

    python tools/pipelined_window_model.py      # self-test against SciPy LAPACK
"""
import numpy as np


def cabs1(z):
    return abs(z.real) + abs(z.imag)


class Shared:
    """A shared-memory array with per-actor access logs."""

    def __init__(self, name, shape, log, dtype=complex):
        self.name, self.a, self.log = name, np.zeros(shape, dtype=dtype), log

    def __getitem__(self, idx):
        self.log.read(self.name, idx)
        return self.a[idx]

    def __setitem__(self, idx, v):
        self.log.write(self.name, idx)
        self.a[idx] = v


class Log:
    """Accesses per actor.  "panel" / "update" are the two sides of an iteration; with the split block
    update the panel side is "panel" (columns 0..3) + "panel4" (last column) and the update side is
    "update" (U, R, A) + "u0a" (block t+1 against pivots 0..3), where "u0a" is ordered after "panel" by a
    flag and runs concurrently with "panel4"."""

    def __init__(self):
        self.actor = None
        self.r = {k: set() for k in ("panel", "panel4", "update", "u0a")}
        self.w = {k: set() for k in ("panel", "panel4", "update", "u0a")}

    def read(self, name, idx):
        if self.actor:
            self.r[self.actor].add((name, idx))

    def write(self, name, idx):
        if self.actor:
            self.w[self.actor].add((name, idx))

    def check_and_reset(self, where):
        def conflict(a, b):
            return (self.w[a] & (self.r[b] | self.w[b])) | (self.w[b] & self.r[a])
        # concurrent pairs: panel (all of it) vs update main; the last panel column vs the early block update
        bad = conflict("panel", "update") | conflict("panel4", "update") | conflict("panel4", "u0a")
        assert not bad, f"race at {where}: {sorted(bad)[:6]}"
        for d in (self.r, self.w):
            for k in d:
                d[k].clear()


def pipelined_solve_T(N, KL, KU, entry, b, P=5, order="panel-first", split_u0=False):
    assert N % P == 0 and (KL + 1) % P == 0
    KV = KL + KU
    RW = KL + P + 1                 # matrix row slots; slot RW = RHS
    NS = RW + 1
    CW = KV + P + 1                 # column slots
    log = Log()
    W = Shared("W", (NS, CW), log)
    stage = Shared("stage", (2, P, CW), log)
    lp = Shared("lp", (2, NS, P), log)
    isp = Shared("isp", (2, NS), log, dtype=bool)
    pivslot = Shared("pivslot", (2, P), log, dtype=int)
    juv = Shared("ju", (2,), log, dtype=int)
    sv = Shared("sv", (N,), log)

    def A_entry(i, c):
        return entry(i, c) if (0 <= i < N and 0 <= c < N and -KL <= c - i <= KU) else 0.0

    def assemble_rows(yI, dst_par):
        # rows 5 yI .. 5 yI + 4, all CW column slots (zeros outside the band)
        for sI in range(P):
            I = P * yI + sI
            for ci in range(CW):
                J = I - KL + ci
                stage[dst_par, sI, J % CW] = A_entry(I, J) if ci <= KV else 0.0

    # ---- prologue (all compute warps) ----
    for i in range(RW):
        for c in range(CW):
            W[i, c % CW] = A_entry(i, c)
    for c in range(N):
        sv[c] = b[c]
    for c in range(CW):
        W[RW, c] = b[c] if c < N else 0.0

    # panel-warp registers
    a = np.zeros((NS, P), dtype=complex)
    lg = np.array([s if s < RW else -10**9 for s in range(NS)])
    pk = np.full(NS, P)
    ju = 0
    L = np.zeros((N, KL), dtype=complex)
    ipiv = np.zeros(N, dtype=np.int32)
    info = [0]

    def panel_side(t, early_u0=None):
        nonlocal ju
        j, par = P * t, t & 1
        log.actor = "panel"
        # ---- load block t: entering rows from the stage, the others from the window ----
        for s_ in range(NS):
            if pk[s_] < P:                          # retired in panel t-1: a new row entered
                for m in range(P):
                    a[s_, m] = stage[par ^ 1, pk[s_], (j + m) % CW]
                lg[s_] = j - P + RW + pk[s_]
                pk[s_] = P
            else:
                for m in range(P):
                    a[s_, m] = W[s_, (j + m) % CW]
        # ---- F(t) ----
        for k in range(P):
            if split_u0 and k == P - 1:
                if early_u0 is not None:
                    early_u0()                       # order "u0a before the last column" (it is concurrent)
                log.actor = "panel4"
            col = j + k
            hi = min(col + KL, N - 1)
            best, bl, bs = -1.0, None, None
            for s in range(RW):
                if pk[s] == P and col <= lg[s] <= hi:
                    mag = cabs1(a[s, k])
                    if mag > best or (mag == best and lg[s] < bl):
                        best, bl, bs = mag, lg[s], s
            jp = bl - col
            ipiv[col] = col + jp + 1
            for s in range(RW):
                if pk[s] == P and lg[s] == col:
                    lg[s] = bl                      # interchange = relabel
            pk[bs], lg[bs] = k, col
            pivslot[par, k] = bs
            piv = a[bs].copy()
            if best == 0.0:
                info[0] = col + 1
                return
            ju = max(ju, min(col + KU + jp, N - 1))
            rinv = 1.0 / piv[k]
            for s in range(NS):
                if pk[s] != P:
                    continue
                l = a[s, k] * rinv
                a[s, k] = l
                if s == RW:
                    sv[col] = l
                elif lg[s] <= hi:
                    L[col, lg[s] - (col + 1)] = l
                for m in range(k + 1, P):
                    a[s, m] -= l * piv[m]
            # multipliers of this column by slot, published column by column (zeros for pivot rows)
            for s in range(NS):
                lp[par, s, k] = a[s, k] if pk[s] == P else 0.0
        for s in range(NS):
            isp[par, s] = pk[s] < P
        juv[par] = ju
        log.actor = None

    def update_side(t, lookahead):
        j, par = P * t, t & 1
        log.actor = None if lookahead else "update"
        if t > 0:
            jo = j - P
            ops = [pivslot[par ^ 1, k] for k in range(P)]
            c_lo, c_hi = (jo + P, jo + 2 * P - 1) if lookahead else (jo + 2 * P, N)
            if lookahead and split_u0:
                # block t already carries pivots 0..P-2 of panel t-1 (u0a): only the last pivot is left, and
                # its row needs no fix-up.  Concurrently: X for the columns past block t.
                for c in range(jo + P, min(jo + 2 * P - 1, N - 1) + 1):
                    cs = c % CW
                    u4 = W[ops[P - 1], cs]
                    for s in range(NS):
                        if isp[par ^ 1, s]:
                            continue
                        W[s, cs] = W[s, cs] - lp[par ^ 1, s, P - 1] * u4
                for c in range(jo + 2 * P, juv[par ^ 1] + 1):
                    cs = c % CW
                    u = [W[ops[m], cs] for m in range(P)]
                    for k in range(1, P):
                        for m in range(k):
                            u[k] -= lp[par ^ 1, ops[k], m] * u[m]
                        W[ops[k], cs] = u[k]
                return
            if lookahead:
                # ---- phase 1: the pivot rows of panel t-1 become rows of U, in place, for every
                # trailing column (one thread per column) ----
                for c in range(jo + P, juv[par ^ 1] + 1):
                    cs = c % CW
                    u = [W[ops[m], cs] for m in range(P)]
                    for k in range(1, P):
                        for m in range(k):
                            u[k] -= lp[par ^ 1, ops[k], m] * u[m]
                        W[ops[k], cs] = u[k]
            # ---- U(t-1): columns c_lo .. c_hi ----
            for c in range(c_lo, min(c_hi, juv[par ^ 1]) + 1):
                cs = c % CW
                u = [W[ops[m], cs] for m in range(P)]
                for s in range(NS):
                    if isp[par ^ 1, s]:
                        continue
                    w = W[s, cs]
                    for m in range(P):
                        w -= lp[par ^ 1, s, m] * u[m]
                    W[s, cs] = w
            if lookahead:
                return
            # ---- R(t-1) ----
            for k in range(P):
                for cs in range(CW):
                    W[ops[k], cs] = stage[par ^ 1, k, cs]
            for m in range(P):
                cs = (jo + m) % CW
                cn = jo + CW + m
                for s in range(RW):
                    if not isp[par ^ 1, s]:
                        W[s, cs] = 0.0
                W[RW, cs] = sv[cn] if cn < N else 0.0
        if lookahead:
            return
        # ---- A(t) ----
        assemble_rows((j + RW) // P, par)
        log.actor = None

    def early_block_update(t):
        """u0a: block t+1 against pivots 0..P-2 of panel t, as soon as the panel warps have published them
        (every thread redoes the small triangular fix-up of the pivot rows for its column)."""
        j, par = P * t, t & 1
        log.actor = "u0a"
        ops = [pivslot[par, k] for k in range(P - 1)]
        for c in range(j + P, min(j + 2 * P - 1, N - 1) + 1):
            cs = c % CW
            u = [W[ops[m], cs] for m in range(P - 1)]
            for k in range(1, P - 1):
                for m in range(k):
                    u[k] -= lp[par, ops[k], m] * u[m]
            for s in range(NS):
                if s in ops:
                    continue
                w = W[s, cs]
                for m in range(P - 1):
                    w -= lp[par, s, m] * u[m]
                W[s, cs] = w
        log.actor = "panel"

    for t in range(N // P):
        update_side(t, True)                      # U0(t-1), then the panel warp is released
        if split_u0:
            # the update side's main work must precede its early block update (same warps); the early
            # update waits for the flag "columns 0..P-2 published"
            if order == "panel-first":
                done = []
                def hook():
                    update_side(t, False); early_block_update(t); done.append(1)
                panel_side(t, early_u0=hook)
            else:
                update_side(t, False)
                panel_side(t, early_u0=lambda: early_block_update(t))
        elif order == "panel-first":
            panel_side(t); update_side(t, False)
        else:
            update_side(t, False); panel_side(t)
        log.check_and_reset(f"panel {t}")
        if info[0]:
            return None, ipiv, L, info[0]
    x = sv.a.copy()
    for j in range(N - 2, -1, -1):
        lm = min(KL, N - 1 - j)
        x[j] -= np.dot(L[j, :lm], x[j + 1:j + 1 + lm])
        l = ipiv[j] - 1
        if l != j:
            x[l], x[j] = x[j], x[l]
    return x, ipiv, L, 0


def _selftest():
    from scipy.linalg import lapack
    rng = np.random.default_rng(11)
    for (N, KL, KU, dom) in [(40, 4, 4, 0.0), (60, 9, 9, 0.0), (120, 14, 14, 2.0), (25, 9, 9, 0.0),
                             (10, 14, 14, 0.0), (5, 4, 4, 0.0), (75, 14, 9, 0.0), (80, 4, 9, 0.0),
                             (150, 24, 24, 0.0)]:
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
        for order, split in (("panel-first", False), ("update-first", False), ("panel-first", True), ("update-first", True)):
            x, ipiv, L, info3 = pipelined_solve_T(N, KL, KU, lambda i, c: A[i, c], b, order=order, split_u0=split)
            assert info3 == 0
            assert np.array_equal(ipiv - 1, piv), (N, KL, KU)
            err = np.abs(x - xr).max() / np.abs(xr).max()
            kv = KL + KU
            Lref = np.array([[lu[kv + i, jj] if jj + i < N else 0 for i in range(1, KL + 1)] for jj in range(N)])
            lerr = np.abs(L - Lref).max()
            assert err < 1e-9 and lerr < 1e-9, (N, KL, KU, order, split, err, lerr)
        print(f"N={N} KL={KL} KU={KU}: pivots identical ({(piv != np.arange(N)).sum()} non-trivial), "
              f"x relerr {err:.2e}, L abs err {lerr:.2e}, no races in either order, with and without the split block update")


if __name__ == "__main__":
    _selftest()
