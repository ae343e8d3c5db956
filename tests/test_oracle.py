"""CPU tests (no GPU): the oracle against the reference's own golden vectors and against
the reference build (oracle/_ref) where present; host-side logic of the product; the
C-ABI library's exported surface."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import parity_common as pc
from oracle import bspline as obs
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EPS = np.finfo(np.float64).eps


def gold(name):
    return json.load(open(os.path.join(GOLD, name)))


def ref_or_skip():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libsuzerain_ref.so not built")
    return ref


# ---------------------------------------------------------------------------
# permutation tables (tests/test_bsmbsm.cpp:48-152)
# ---------------------------------------------------------------------------
def test_permutation_golden():
    g = gold("bsmbsm_perm.json")
    S, n = g["S"], g["n"]
    import suzerain_b200 as sz
    L = sz.lib.load()
    for i, v in g["q"].items():
        assert port.q(S, n, int(i)) == v
        assert L.szb_bsmbsm_q(S, n, int(i)) == v
    for i, v in g["qinv"].items():
        assert port.qinv(S, n, int(i)) == v
        assert L.szb_bsmbsm_qinv(S, n, int(i)) == v
    # identity_relation (:170-184)
    for S in range(1, 6):
        for n in range(1, 10):
            for i in range(S * n):
                assert port.qinv(S, n, port.q(S, n, i)) == i
                assert L.szb_bsmbsm_qinv(S, n, L.szb_bsmbsm_q(S, n, i)) == i


def test_bsmbsm_construct_matches_reference_formulas():
    import suzerain_b200 as sz
    L = sz.lib.load()
    for S, n, kl, ku in [(3, 10, 4, 4), (5, 96, 6, 6), (5, 24, 4, 4), (1, 7, 2, 3)]:
        A, want = L.szb_bsmbsm_construct(S, n, kl, ku), port.bsmbsm(S, n, kl, ku)
        for k, v in want.items():
            assert getattr(A, k) == v
    # suzerain/gbmatrix.h:51-68
    assert L.szb_gbmatrix_offset(9, 4, 4, 3, 5) == port.gb_offset(9, 4, 4, 3, 5)
    assert L.szb_gbmatrix_in_band(4, 4, 9, 5) == 1 and L.szb_gbmatrix_in_band(4, 4, 10, 5) == 0


# ---------------------------------------------------------------------------
# 30 x 30 BSMBSM solve with a 30-digit Mathematica answer (tests/test_bsmbsm.cpp:698-966)
# ---------------------------------------------------------------------------
def golden_system(cplx=False):
    g = gold("bsmbsm_solve.json")
    A = port.bsmbsm(g["S"], g["n"], g["kl"], g["ku"])
    band = {k: np.array(g[k]).reshape(A["n"], A["ld"]) for k in ("M", "D1", "D2")}
    papt = np.full((A["N"], A["LD"] + A["KL"]), np.nan, dtype=np.complex128)
    view = papt[:, A["KL"]:]                                  # packf: offset by KL rows
    s = 1j if cplx else 1.0
    for (i, j, name, alpha) in [(0, 0, "M", 1.0), (1, 1, "M", 2.0), (2, 2, "M", 4.0),
                                (0, 1, "D1", 1 / 5), (2, 1, "D1", 1 / 10),
                                (1, 0, "D2", 1 / 7), (1, 2, "D2", 1 / 14)]:
        port.zpack(A, i, j, s * alpha, band[name], view, A["LD"])
    port.zpack(A, 0, 2, 0.0, None, view, A["LD"])
    port.zpack(A, 2, 0, 0.0, None, view, A["LD"])
    return g, A, papt


@pytest.mark.parametrize("cplx", [False, True])
def test_golden_solve_port(cplx):
    g, A, papt = golden_system(cplx)
    N, KL, KU = A["N"], A["KL"], A["KU"]
    assert not np.isnan(papt[:, KL:][np.abs(np.arange(A["LD"])[None, :] - KU
                                           + np.arange(N)[:, None] - N // 2) < 0]).any()
    ab = np.where(np.isnan(papt), 0, papt).T.copy()           # (2KL+KU+1, N)
    BR = np.array(g["BR"], dtype=np.float64)
    b = port.aPxpby("N", A["S"], A["n"], 1.0, (-BR + 1j * BR) if cplx else BR.astype(complex), 0.0, None)
    afb, ipiv, info = port.zgbtf2(ab, N, KL, KU)
    assert info == 0
    x = port.aPxpby("T", A["S"], A["n"], 1.0, port.zgbtrs("N", afb, N, KL, KU, ipiv, b), 0.0, None)
    XR = np.array(g["XR"])
    want = (1 + 1j) * XR if cplx else XR
    assert np.abs(x - want).max() <= 1.8e4 * EPS * np.abs(XR).max()
    # LAPACK itself (SciPy's OpenBLAS) picks the same pivots on this system
    from scipy.linalg import lapack
    lu, piv, info = lapack.zgbtrf(ab, KL, KU)
    assert np.array_equal(piv + 1, ipiv)


# ---------------------------------------------------------------------------
# B-spline collocation operators (tests/test_bsplineop.cpp:70-742)
# ---------------------------------------------------------------------------
def band_to_dense_T(band, m, n, kl, ku, ld):
    """Golden matrices are D^T in column-major general band storage."""
    Dt = np.zeros((m, n))
    b = np.array(band).reshape(n, ld)
    for j in range(n):
        for i in range(max(0, j - ku), min(m, j + kl + 1)):
            Dt[i, j] = b[j, ku + i - j]
    return Dt


@pytest.mark.parametrize("impl", ["oracle", "cabi"])
def test_bsplineop_collocation_golden(impl):
    import suzerain_b200 as sz
    g = gold("bsplineop_colloc.json")
    for case in g["cases"]:
        k, bp = case["k"], np.array(case["breakpoints"])
        nderiv = max(d["d"] for d in case["D_T"])
        op = (obs.make_bsplineop(k, bp, nderiv=max(nderiv, 1)) if impl == "oracle"
              else sz.BsplineOp.from_breakpoints(k, bp, nderiv=max(nderiv, 1)))
        for d in case["D_T"]:
            want = band_to_dense_T(d["band"], d["m"], d["n"], d["kl"], d["ku"], d["ld"]).T
            got = op.dense(d["d"])
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 1000 * EPS * max(1.0, np.abs(want).max()), (case["name"], d["d"])
            # CHECK_GBMATRIX_CLOSE compares across different bandwidths: the golden's are an
            # upper bound on what the exact-zero scan (bsplineop.c:403-517) leaves
            assert int(op.kl[d["d"]]) <= d["kl"] and int(op.ku[d["d"]]) <= d["ku"], (case["name"], d["d"])
    # k = 8 on a single interval: operator actions (:697-742)
    s = g["septic"]
    op = (obs.make_bsplineop(8, np.array(s["breakpoints"]), nderiv=2) if impl == "oracle"
          else sz.BsplineOp.from_breakpoints(8, np.array(s["breakpoints"]), nderiv=2))
    for act in s["actions"]:
        x = np.array(act["x"]).reshape(act["nrhs"], -1)
        y = act["alpha"] * (op.dense(act["d"]) @ x.T).T
        want = np.array(act["y"]).reshape(act["nrhs"], -1)
        assert np.abs(y - want).max() <= 1000 * EPS * np.abs(want).max()


def test_cabi_bsplineop_matches_oracle_on_stretched_grids():
    import suzerain_b200 as sz
    for (n, k, htdelta) in [(24, 6, 2.0), (96, 8, 3.0), (40, 4, 0.0), (48, 10, -2.0), (33, 5, 1.0)]:
        bp = obs.breakpoints(n, k, 0.0, 2.0, htdelta)
        assert np.allclose(bp, sz.htstretch_breakpoints(n, k, 0.0, 2.0, htdelta), rtol=0, atol=4 * EPS)
        a, b = obs.make_bsplineop(k, bp), sz.BsplineOp.from_breakpoints(k, bp)
        assert (a.n, a.ld, a.max_kl, a.max_ku) == (b.n, b.ld, b.max_kl, b.max_ku)
        assert np.array_equal(a.kl, b.kl) and np.array_equal(a.ku, b.ku)
        for d in range(3):
            scale = np.abs(a.storage[d]).max()
            assert np.abs(a.storage[d] - b.storage[d]).max() <= 2e-13 * scale
        assert np.abs(a.greville - b.greville()).max() <= 8 * EPS


# ---------------------------------------------------------------------------
# numpy port vs the reference's own object code
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,kw", [("tiny_16x24x16", dict(max_pencils=5)),
                                    ("bl_1024x256x512", dict(max_pencils=4, Ny=32)),
                                    ("tiny_16x24x16", dict(max_pencils=4, k=8, Ny=30))])
def test_port_matches_reference_build(cfg, kw):
    ref = ref_or_skip()
    case = pc.make_case(cfg, **kw)
    Pr = ref.Problem(case.bop, case.scenario, case.refs, case.bc_dict(), case.nrbc)
    Pp = port.Problem(case.bop, case.scenario, case.refs, case.bc_dict(), case.nrbc)
    x = case.x.reshape(len(case.km), -1)
    for packf in (False, True):
        for with_bc in (False, True):
            a = Pr.assemble(case.phi, float(case.km[1]), float(case.kn[1]), packf=packf, with_bc=with_bc)
            b = Pp.assemble(case.phi, float(case.km[1]), float(case.kn[1]), packf=packf, with_bc=with_bc)
            assert np.array_equal(np.isnan(a), np.isnan(b))           # NaN-poisoned padding untouched
            m = ~np.isnan(a)
            assert pc.relmax(b[m], a[m]) <= 1e-14
    beta = 0.3 - 0.1j
    assert pc.relmax(Pp.accumulate(case.phi, case.km, case.kn, x, beta=beta, y=x),
                     Pr.accumulate(case.phi, case.km, case.kn, x, beta=beta, y=x)) <= 1e-13
    for solver in ("zgbsv", "zcgbsvx"):
        ra = Pr.invert(solver, case.phi, case.km, case.kn, x, want_ipiv=True)
        rb = Pp.invert(solver, case.phi, case.km, case.kn, x, want_ipiv=True)
        assert ra["info"] == 0 and rb["info"] == 0
        assert np.array_equal(ra["ipiv"], rb["ipiv"])
        assert pc.relmax(rb["x"], ra["x"]) <= 1e-12


def test_reference_apply_then_solve_is_identity():
    """The reference's own consistency check (tests/test_rholut_imexop.cpp:86-302) replayed on
    the reference build: (M + phi L) applied to x, then solved without boundary rows, returns x."""
    ref = ref_or_skip()
    case = pc.make_case("tiny_16x24x16", max_pencils=4, phi=complex(-0.01, 0.002))
    nobc = dict(case.bc_dict(), enforce_lower=0, enforce_upper=0)
    P = ref.Problem(case.bop, case.scenario, case.refs, nobc, None)
    x = case.x.reshape(len(case.km), -1)
    y = P.accumulate(case.phi, case.km, case.kn, x)
    r = P.invert("zgbsv", case.phi, case.km, case.kn, y)
    assert r["info"] == 0
    assert pc.relmax(r["x"], x) <= 1e-10


def test_reference_pack_variants_agree():
    """packf == packc up to the KL-row offset (tests/test_rholut_imexop.cpp:215-230)."""
    ref = ref_or_skip()
    case = pc.make_case("tiny_16x24x16", max_pencils=3)
    P = ref.Problem(case.bop, case.scenario, case.refs, case.bc_dict(), None)
    c = P.assemble(case.phi, float(case.km[2]), float(case.kn[2]), packf=False, with_bc=False)
    f = P.assemble(case.phi, float(case.km[2]), float(case.kn[2]), packf=True, with_bc=False)
    assert np.isnan(f[:, :P.KL]).all()
    assert np.array_equal(np.nan_to_num(f[:, P.KL:], nan=7.0), np.nan_to_num(c, nan=7.0))


# ---------------------------------------------------------------------------
# host-side logic of the product (no GPU)
# ---------------------------------------------------------------------------
def test_wavegrid_wavenumbers_and_dealiasing():
    import suzerain_b200 as sz
    Nx, Nz, Lx, Lz = 16, 12, 4 * np.pi, 4 * np.pi / 3
    g = sz.wavegrid(Nx, Nz, Lx, Lz)
    km, kn, act = sz.wavenumbers(g)
    nx = g.dkex - g.dkbx
    assert len(km) == nx * (g.dkez - g.dkbz) == (g.dNx // 2 + 1) * g.dNz
    wn = lambda N, i: i if i < N // 2 + 1 else -N + i              # suzerain/inorder.h:92-96
    for p in range(len(km)):
        m, n = p % nx, p // nx
        wm, wz = wn(g.dNx, m), wn(g.dNz, n)
        assert km[p] == (2 * np.pi / Lx) * wm and kn[p] == (2 * np.pi / Lz) * wz
        assert act[p] == (abs(wm) <= (Nx - 1) // 2 and abs(wz) <= (Nz - 1) // 2)
    assert act.sum() == (Nx // 2) * (Nz - 1)                          # active (kx,kz)
    L = sz.lib.load()
    assert L.szb_wavegrid_nactive(C.byref(g)) == act.sum()


def test_solver_spec_grammar():
    import suzerain_b200 as sz
    d = sz.SolverSpec()
    assert (d.method, d.aiter, d.diter, d.tolsc) == ("zcgbsvx", 1, 5, 0.0)   # specification_zgbsv.cpp:46-54
    c = sz.lib.load().szb_zgbsv_spec_default()
    assert (c.method, c.aiter, c.diter, c.tolsc) == (1, 1, 5, 0.0)
    assert sz.SolverSpec.parse("zgbsv").method == "zgbsv"
    s = sz.SolverSpec.parse("zcgbsvx,reuse=false,aiter=2,siter=-1,diter=7,tolsc=0.5")
    assert (s.aiter, s.diter, s.tolsc) == (2, 7, 0.5)
    c = sz.lib.load().szb_zgbsv_spec_default()
    assert (c.equil, c.reuse, c.siter) == (0, 0, -1)
    # the six specifications of apps/perfect/test_implicit_solvers.sh:24-30
    for text, want in (("zgbsv", ("zgbsv", False, False, 1, -1, 5, 0.0)),
                       ("zgbsvx,equil=false", ("zgbsvx", False, False, 1, -1, 5, 0.0)),
                       ("zgbsvx,equil=true", ("zgbsvx", True, False, 1, -1, 5, 0.0)),
                       ("zcgbsvx,reuse=false,aiter=1,siter=-1,diter=5,tolsc=0", ("zcgbsvx", False, False, 1, -1, 5, 0.0)),
                       ("zcgbsvx,reuse=true,aiter=1,siter=-1,diter=5,tolsc=0", ("zcgbsvx", False, True, 1, -1, 5, 0.0)),
                       ("ZCGBSVX, reuse=TRUE, aiter=5, siter=25, diter=5, tolsc=0", ("zcgbsvx", False, True, 5, 25, 5, 0.0)),
                       ("", ("zcgbsvx", False, False, 1, -1, 5, 0.0))):
        s = sz.SolverSpec.parse(text)
        assert (s.method, s.equil, s.reuse, s.aiter, s.siter, s.diter, s.tolsc) == want, text
        cs = s.c()
        assert (cs.method, cs.equil, cs.reuse, cs.siter) == ({"zgbsv": 0, "zcgbsvx": 1, "zgbsvx": 2}[want[0]],
                                                             int(want[1]), int(want[2]), want[4])
    assert sz.SolverSpec.parse("zgbsv").in_place() and not sz.SolverSpec.parse("zgbsvx").in_place()
    for bad in ("zgbsv,equil=true", "zgbsvx,reuse=true", "zcgbsvx,equil=true", "zcgbsvx,aiter", "zcgbsvx,aiter=1,aiter=2",
                "dgbsv", "zgbsvx,", "zcgbsvx,reuse=maybe"):
        with pytest.raises(ValueError):
            sz.SolverSpec.parse(bad)


def test_shard_bounds_balance_active_pencils():
    import suzerain_b200 as sz
    from suzerain_b200 import shard, synth
    g = sz.wavegrid(192, 192, synth.LX, synth.LZ)
    w = shard.active_rows(g)
    assert w.sum() == 96 * 191
    for world in (1, 2, 3, 4, 8):
        b = shard.shard_bounds(w, world)
        assert b[0] == 0 and b[-1] == len(w) and all(b[i] < b[i + 1] for i in range(world))
        loads = [w[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(loads) <= 1.06 * w.sum() / world
        subs = [shard.shard_wavegrid(g, r, world) for r in range(world)]
        assert sum(sz.lib.load().szb_wavegrid_nactive(C.byref(s)) for s in subs) == w.sum()
        assert shard.owner_of_zero_zero(g, world) == 0


def test_window_lu_models_match_lapack():
    """The executable models of the fused kernels' index algebra and schedules (tools/): the pipelined one race-checks
    v4's split block update; the synchronous one runs v5's three phases with their concurrent roles (speculative panel,
    exact fallback, tail rows of the trailing update deferred to the next phase 1), checks that no role writes what
    another role of the same phase touches, that the barrier between the deferred tail rows and the assembly is needed,
    and that pivots and solutions are LAPACK's."""
    import importlib.util
    for name in ("window_lu_model", "blocked_window_model", "pipelined_window_model", "sync_window_model"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod._selftest()


# ---------------------------------------------------------------------------
# C ABI surface
# ---------------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol():
    import suzerain_b200 as sz
    hdr = open(os.path.join(ROOT, "include", "suzerain_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(szb_[a-zA-Z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 40
    lib = C.CDLL(sz.lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes prototype table covers the same surface
    assert sorted(sz.lib.PROTOTYPES) == names
    assert sz.lib.load().szb_version().startswith(b"suzerain_b200")


def test_cabi_argument_checks_follow_lapack_convention():
    """Bad arguments come back as -k (suzerain/blas_et_al/blas.c:68-74) before any CUDA call."""
    import suzerain_b200 as sz
    L = sz.lib.load()
    two = (C.c_double * 2)(1.0, 0.0)
    assert L.szb_zgbtrf_batch(-1, 1, 1, None, 4, 0, None, None, 1, None) == -1
    assert L.szb_zgbtrf_batch(4, 1, 1, None, 4, 0, None, None, 1, None) == -4
    assert L.szb_zgbtrs_batch(b"X", 4, 1, 1, 1, None, 4, 0, None, None, 4, 0, 1, None) == -1
    assert L.szb_bsmbsm_zaPxpby_batch(b"N", 5, 4, two, None, two, None, 1, None) == -5
    assert L.szb_imexop_accumulate_batch(None, two, 1, None, None, None, None, 0, 0, two, None, 0, 0, None) == -1
    assert L.szb_imexop_invert_batch(None, None, two, 1, None, None, None, None, 0, 0, 0, None, None, None,
                                     None, None) == -1
    out = C.c_void_p()
    assert L.szb_bsplineop_alloc(0, 4, None, 2, C.byref(out)) < 0


def test_reference_diffwave_matches_formula():
    """oracle/_ref now also carries the reference's suzerain/diffwave.c (built unmodified): check the
    binding and the stub gsl_sf_pow_int against the closed form alpha (i kx)^dx (i kz)^dz x with
    dealiased / Nyquist modes zeroed (suzerain/diffwave.c:65-198, inorder.h:282-293)."""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    Nx, dNx, Nz, dNz, Ny = 8, 12, 6, 9, 5
    Lx, Lz = 4 * np.pi, 2.0
    grid = (Nx, dNx, 0, dNx // 2 + 1, Nz, dNz, 0, dNz)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((dNz, dNx // 2 + 1, Ny)) + 1j * rng.standard_normal((dNz, dNx // 2 + 1, Ny))
    y = rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape)
    alpha, beta = 0.7 - 0.3j, 1.5 + 0.25j

    def freq(N, dN, i):                          # suzerain_inorder_wavenumber_diff
        return i if i < (N + 1) // 2 else (-dN + i if i >= dN - (N - 1) // 2 else 0)

    def absw(dN, i):
        return i if i < dN // 2 + 1 else dN - i

    for dx, dz in [(0, 0), (1, 0), (0, 2), (2, 1)]:
        want = np.zeros_like(x)
        for n in range(dNz):
            keepn = freq(Nz, dNz, n) != 0 if dz > 0 else absw(dNz, n) <= (Nz - 1) // 2
            for m in range(dNx // 2 + 1):
                keepm = freq(Nx, dNx, m) != 0 if dx > 0 else absw(dNx, m) <= (Nx - 1) // 2
                if keepn and keepm:
                    kx, kz = 2 * np.pi / Lx * freq(Nx, dNx, m), 2 * np.pi / Lz * freq(Nz, dNz, n)
                    want[n, m] = alpha * (1j * kx) ** dx * (1j * kz) ** dz * x[n, m]
        got = oref.diffwave(dx, dz, alpha, x, Lx, Lz, grid)
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
        got = oref.diffwave(dx, dz, alpha, x, Lx, Lz, grid, beta=beta, y=y)
        assert np.abs(got - (want + beta * y)).max() <= 1e-13 * np.abs(want + beta * y).max()


def test_reference_solver_specifications_agree_like_its_own_test():
    """apps/perfect/test_implicit_solvers.sh:24-50 on the tiny synthetic case: the six --solver
    specifications, run through the reference's own zgbsvx / zcgbsvx with its solver-state chain
    (oracle/_ref, ref_invert_spec_batch), agree with plain zgbsv."""
    pytest.importorskip("scipy")
    import parity_common as pc
    case = pc.make_case("tiny_16x24x16")
    try:
        P = pc.oracle_problem(case, "ref")
    except Exception as e:                                   # noqa: BLE001
        pytest.skip(f"oracle/_ref not built: {e}")
    npen = len(case.km)
    x = case.x.reshape(npen, -1)
    base = P.invert("zgbsv", case.phi, case.km, case.kn, x)["x"]
    for kw in (dict(method="zgbsv"), dict(method="zgbsvx", equil=False), dict(method="zgbsvx", equil=True),
               dict(method="zcgbsvx"), dict(method="zcgbsvx", reuse=True, rowlen=8),
               dict(method="zcgbsvx", reuse=True, aiter=5, siter=25, rowlen=8)):
        r = P.invert_spec(case.phi, case.km, case.kn, x, **kw)
        assert r["info"] == 0, kw
        assert pc.relmax(r["x"], base) <= 1e-13, kw
    # larger |phi| makes zlaqgb equilibrate: rows at 100x, rows and columns at 1000x
    for scale, code in ((100, 1), (1000, 3)):
        r = P.invert_spec(case.phi * scale, case.km, case.kn, x, method="zgbsvx", equil=True)
        assert r["info"] == 0 and set(r["stats"][:, 0].astype(int)) == {code}


def test_public_headers_are_plain_c():
    """The drop-in boundary is a C ABI: both headers must compile as C99 and as C++11 on their own."""
    import shutil
    import subprocess
    import tempfile
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = '#include "suzerain_b200.h"\n#include "suzerain_b200_fft.h"\nint main(void) { szb_zgbsv_spec s = szb_zgbsv_spec_default(); (void) s; return 0; }\n'
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "hdr.c")
        open(p, "w").write(src)
        inc = os.path.join(ROOT, "include")
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I" + inc, p], check=True)
        if shutil.which("g++"):
            subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I" + inc, "-x", "c++", p], check=True)


def test_reference_bsplineop_family_agrees_with_the_dense_operators():
    """oracle/ref_shim/ref_glue.c restates the four suzerain_bsplineop_{accumulate,apply}{,_complex} drivers
    (suzerain/bsplineop.c:222-381) around the reference's own dgbmv / zgbmv_d_z: all of them against the dense
    collocation operators, incl. the in-place forms and beta != 0."""
    case = pc.make_case("tiny_16x24x16", max_pencils=4, k=6, Ny=41)
    P = pc.oracle_problem(case, "ref")
    rng = np.random.default_rng(3)
    x = rng.standard_normal((7, 41)); y = rng.standard_normal((7, 41))
    xc = x + 1j * rng.standard_normal((7, 41)); yc = y - 2j * x
    for d in range(3):
        D = case.bop.dense(d)
        scale = np.abs(D).sum(axis=1).max()
        assert np.abs(P.bsplineop_accumulate(d, 1.3, x, -0.6, y) - (1.3 * x @ D.T - 0.6 * y)).max() <= 1e-14 * scale * 10
        assert np.abs(P.bsplineop_accumulate(d, 1.3, x) - 1.3 * x @ D.T).max() <= 1e-14 * scale * 10
        assert np.abs(P.bsplineop_apply(d, 2.0, x) - 2.0 * x @ D.T).max() <= 1e-14 * scale * 10
        assert np.abs(P.bsplineop_apply(d, 2.0, xc) - 2.0 * xc @ D.T).max() <= 1e-14 * scale * 10
        want = (1.3 + 0.4j) * xc @ D.T + (0.7 - 0.2j) * yc
        assert np.abs(P.bsplineop_accumulate_complex(d, 1.3 + 0.4j, xc, 0.7 - 0.2j, yc) - want).max() <= 1e-14 * scale * 10


def test_equation_of_state_port_matches_the_reference_known_answers():
    """oracle.port.p_T_mu_lambda against the 200-bit Sage values of tests/test_rholut.cpp:155-240 (fixture:
    tests/golden/rholut_known_answers.json), and the three reference coefficients of collect_references against
    the closed forms the reference's own tests check them with (tests/test_rholut.cpp:395-420, 845-860)."""
    g = json.load(open(os.path.join(pc.ROOT, "tests", "golden", "rholut_known_answers.json")))
    p, T, mu, lam = port.p_T_mu_lambda(g["alpha"], g["beta"], g["gamma"], g["Ma"], g["rho"], np.array(g["m"]), g["e"])
    eps = np.finfo(float).eps
    for got, key in ((p, "p"), (T, "T"), (mu, "mu"), (lam, "lambda")):
        assert abs(got / g[key] - 1) <= 10 * eps                     # BOOST_CHECK_CLOSE(.., eps * 1e3) is in percent
    gamma, Ma, mu, rho, pp = 1.4, 3.5, 4181.0, 67.0, 55.0
    m = np.array([144.0, 233.0, 377.0])
    e = pp / (gamma - 1) + Ma * Ma * (m @ m) / (2 * rho)
    assert abs(port.explicit_mu_div_grad_T_refcoeff_div_grad_rho(gamma, mu, rho, e, pp)
               / (mu / rho / rho * ((gamma - 1) * e - 2 * pp)) - 1) <= 10 * eps
    assert abs(port.explicit_div_e_plus_p_u_refcoeff_div_m(rho, e, pp) / ((e + pp) / rho) - 1) <= 10 * eps
    assert np.allclose(port.explicit_div_e_plus_p_u_refcoeff_grad_rho(gamma, rho, m, e, pp),
                       m * ((gamma - 2) * e - 2 * pp) / rho ** 2, rtol=10 * eps, atol=0)


def test_collect_references_port_against_a_point_by_point_loop():
    """The vectorised oracle against the reference's loop written out point by point with Kahan sums
    (apps/perfect/perfect.cpp:1279-1393, perfect.hpp:78-86), incl. the inviscid top plane of one-sided grids."""
    import math
    rng = np.random.default_rng(8)
    Ny, Nz, Nx = 4, 3, 5
    gamma, Ma, alpha, beta = 1.4, 1.5, 0.0, 2.0 / 3.0
    rho = 1 + 0.2 * rng.uniform(-1, 1, (Ny, Nz, Nx)); u = 0.3 * rng.standard_normal((3, Ny, Nz, Nx))
    T = 1 + 0.2 * rng.uniform(-1, 1, (Ny, Nz, Nx))
    p = rho * T / gamma
    e = p / (gamma - 1) + Ma * Ma * rho * (u ** 2).sum(axis=0) / 2
    s = np.stack([e, rho * u[0], rho * u[1], rho * u[2], rho])
    for top in (False, True):
        got = port.collect_references(alpha, beta, gamma, Ma, s, top)
        want = np.zeros((42, Ny))
        for j in range(Ny):
            rows = [[] for _ in range(42)]
            for k in range(Nz):
                for i in range(Nx):
                    ee, mx, my, mz, r = s[:, j, k, i]
                    pr = (gamma - 1) * (ee - Ma * Ma / r * (mx * mx + my * my + mz * mz) / 2)
                    TT = gamma * pr / r
                    mu = TT ** beta * (0 if (top and j == Ny - 1) else 1)
                    ux, uy, uz = mx / r, my / r, mz / r
                    u2 = ux * ux + uy * uy + uz * uz
                    nu = mu / r
                    cg = ((gamma - 2) * ee - 2 * pr) / (r * r)
                    vals = [r, pr, pr * pr, TT, math.sqrt(TT), ux, uy, uz, u2, ux * ux, ux * uy, ux * uz, uy * uy, uy * uz,
                            uz * uz, nu, nu * ux, nu * uy, nu * uz, nu * u2, nu * ux * ux, nu * ux * uy, nu * ux * uz,
                            nu * uy * uy, nu * uy * uz, nu * uz * uz, cg * mx, cg * my, cg * mz, (ee + pr) / r,
                            mu / (r * r) * ((gamma - 1) * ee - 2 * pr), mx, my, mz, ee, mx * mx / r, mx * my / r,
                            mx * mz / r, my * my / r, my * mz / r, mz * mz / r, ee * ee / r]
                    for q, v in enumerate(vals):
                        rows[q].append(v)
            want[:, j] = [math.fsum(r_) for r_ in rows]
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
        assert len(port.REFERENCE_QUANTITIES) == 42 and port.REFERENCE_QUANTITIES[5:31][0] == "u"
