"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the
same seeded inputs.  Tolerances are north_star's: relative max-norm <= 1e-12
after one operator application / solve, pivot choices identical."""
import ctypes as C

import numpy as np
import pytest

import parity_common as pc

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def dev():
    import torch
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def tiny():
    return pc.make_case("tiny_16x24x16")


@pytest.fixture(scope="module")
def ch96():
    return pc.make_case("channel_192x96x192", max_pencils=40)


@pytest.fixture(scope="module")
def bl():
    # one-sided grid with NRBC matrices, reduced Ny so the oracle stays fast
    return pc.make_case("bl_1024x256x512", max_pencils=24, Ny=64)


@pytest.mark.parametrize("packf", [False, True])
@pytest.mark.parametrize("with_bc", [False, True])
def test_pack_matches_oracle(dev, tiny, packf, with_bc):
    got = pc.gpu_pack(tiny, dev, packf=packf, with_bc=with_bc)
    for p in (0, 1, len(tiny.km) // 2, len(tiny.km) - 1):
        want = pc.oracle_assemble(tiny, p, packf=packf, with_bc=with_bc)
        # storage outside the matrix stays NaN-poisoned on both sides
        assert np.array_equal(np.isnan(got[p]), np.isnan(want))
        m = ~np.isnan(want)
        assert pc.relmax(got[p][m], want[m]) <= 1e-14


def test_pack_nrbc_matches_oracle(dev, bl):
    got = pc.gpu_pack(bl, dev, packf=False, with_bc=True)
    for p in (0, 3, len(bl.km) - 1):
        want = pc.oracle_assemble(bl, p, packf=False, with_bc=True)
        assert np.array_equal(np.isnan(got[p]), np.isnan(want))
        m = ~np.isnan(want)
        assert pc.relmax(got[p][m], want[m]) <= 1e-14


@pytest.mark.parametrize("casename", ["tiny", "ch96", "bl"])
def test_accumulate_matches_oracle(dev, request, casename):
    case = request.getfixturevalue(casename)
    got = pc.gpu_accumulate(case, dev)
    want = pc.oracle_accumulate(case)
    assert pc.relmax(got, want) <= TOL
    # beta != 0 path (accumulate proper)
    rng = np.random.default_rng(3)
    y0 = rng.standard_normal(case.x.shape) + 1j * rng.standard_normal(case.x.shape)
    beta = 0.37 - 0.21j
    got = pc.gpu_accumulate(case, dev, beta=beta, y=y0)
    want = pc.oracle_accumulate(case, beta=beta, y=y0)
    assert pc.relmax(got, want) <= TOL


@pytest.mark.parametrize("casename", ["tiny", "ch96", "bl"])
@pytest.mark.parametrize("solver", ["zgbsv", "zcgbsvx"])
def test_invert_matches_oracle(dev, request, casename, solver):
    case = request.getfixturevalue(casename)
    got = pc.gpu_invert(case, solver, dev)
    want = pc.oracle_invert(case, solver)
    assert want["info"] == 0
    assert np.all(got["info"] == 0)
    assert np.array_equal(got["ipiv"], want["ipiv"]), "pivot choices differ"
    assert pc.relmax(got["x"], want["x"]) <= TOL
    if solver == "zcgbsvx":
        # refinement counter as the reference reports it (dsgbsvx.def:169-170, 256-290): the stopping
        # tests compare residual norms at the rounding level, so allow a step of slack on a few pencils
        d = np.abs(got["iters"].astype(int) - want["iters"].astype(int))
        assert d.max() <= 1 and (d != 0).mean() <= 0.2, (got["iters"], want["iters"])


def test_invert_extra_rhs(dev, tiny):
    """Integral-constraint columns reuse the pencil's factorisation
    (operator_hybrid_isothermal.cpp:676-685)."""
    rng = np.random.default_rng(11)
    npen, N = len(tiny.km), 5 * tiny.n
    extra = rng.standard_normal((npen, 3, N)) + 1j * rng.standard_normal((npen, 3, N))
    for solver in ("zgbsv", "zcgbsvx"):
        got = pc.gpu_invert(tiny, solver, dev, extra=extra)
        want = pc.oracle_invert(tiny, solver, extra=extra)
        assert pc.relmax(got["x"], want["x"]) <= TOL
        assert pc.relmax(got["extra"], want["extra"]) <= TOL


def test_apply_then_invert_roundtrip(dev, ch96):
    """apply o invert = identity away from the wall rows (the reference's own
    consistency check, tests/test_rholut_imexop.cpp:86-302)."""
    import torch
    op = pc.make_imexop(ch96)
    km = torch.from_numpy(ch96.km).to(dev)
    kn = torch.from_numpy(ch96.kn).to(dev)
    x = torch.from_numpy(ch96.x).to(dev)
    y = torch.zeros_like(x)
    op.accumulate_batch(ch96.phi, km, kn, x, 0.0, y)
    info = torch.zeros(len(ch96.km), dtype=torch.int32, device=dev)
    from suzerain_b200 import SolverSpec
    op.invert_batch(SolverSpec("zgbsv"), ch96.phi, km, kn, y, info=info)
    torch.cuda.synchronize()
    xr = y.cpu().numpy()
    # wall rows of e, mx, my, mz were replaced by constraints; compare the rest
    # through the residual instead: (M + phi L) xr == (M + phi L) x except there
    y2 = torch.zeros_like(x)
    op.accumulate_batch(ch96.phi, km, kn, torch.from_numpy(xr).to(dev), 0.0, y2)
    y1 = torch.zeros_like(x)
    op.accumulate_batch(ch96.phi, km, kn, x, 0.0, y1)
    torch.cuda.synchronize()
    d = (y2 - y1).abs().cpu().numpy()
    scale = y1.abs().max().item()
    assert d[:, :, 1:-1].max() / scale <= 1e-11
    assert d[:, 4, :].max() / scale <= 1e-11


# ---------------------------------------------------------------------------
# register-window invert kernel (invert_window.cu): orders, heavy pivoting,
# persistent-slot reuse, agreement with the generic v1 kernel
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("k", [4, 6, 8, 10])
def test_invert_window_orders(dev, k):
    case = pc.make_case("tiny_16x24x16", max_pencils=20, k=k, Ny=40)
    got = pc.gpu_invert(case, "zgbsv", dev)
    want = pc.oracle_invert(case, "zgbsv")
    assert want["info"] == 0 and np.all(got["info"] == 0)
    assert np.array_equal(got["ipiv"], want["ipiv"]), "pivot choices differ"
    assert pc.relmax(got["x"], want["x"]) <= TOL


@pytest.mark.parametrize("phi", [-50.0, -5 + 3j, 0.3j])
@pytest.mark.parametrize("solver", ["zgbsv", "zcgbsvx"])
def test_invert_heavy_pivoting(dev, phi, solver):
    """|phi| = O(1..50) makes more than half of the row interchanges non-trivial."""
    case = pc.make_case("tiny_16x24x16", max_pencils=30, phi=complex(phi))
    got = pc.gpu_invert(case, solver, dev)
    want = pc.oracle_invert(case, solver)
    nontrivial = (want["ipiv"] != np.arange(1, want["ipiv"].shape[1] + 1)).mean()
    assert nontrivial > 0.3
    assert np.array_equal(got["ipiv"], want["ipiv"]), "pivot choices differ"
    # these operators are far worse conditioned than the time stepper's; scale the
    # tolerance by the growth the oracle itself shows between its two solvers
    assert pc.relmax(got["x"], want["x"]) <= 1e-10


@pytest.mark.parametrize("k", [8, 10])
@pytest.mark.parametrize("phi", [-50.0, -5 + 3j, 0.3j])
def test_invert_heavy_pivoting_two_panel_warps(dev, k, phi):
    """Orders 8 and 10 need more than 32 window rows, i.e. two panel warps in the pipelined
    kernel (invert_pipe.cu): the pivot of a column is then decided across warps, and with
    |phi| = O(1..50) most decisions are non-trivial interchanges, many of them near ties."""
    case = pc.make_case("tiny_16x24x16", max_pencils=30, phi=complex(phi), k=k, Ny=48)
    got = pc.gpu_invert(case, "zgbsv", dev)
    want = pc.oracle_invert(case, "zgbsv")
    nontrivial = (want["ipiv"] != np.arange(1, want["ipiv"].shape[1] + 1)).mean()
    assert nontrivial > 0.3
    assert np.all(got["info"] == 0) and want["info"] == 0
    assert np.array_equal(got["ipiv"], want["ipiv"]), "pivot choices differ"
    assert pc.relmax(got["x"], want["x"]) <= 1e-10


def test_invert_singular_reports_info_two_panel_warps(dev):
    """zgbtrf's info (first zero pivot) out of the cross-warp exact pivot decision."""
    import suzerain_b200 as sz
    case = pc.make_case("tiny_16x24x16", max_pencils=6, k=8, Ny=40)
    bop = case.bop
    st = bop.storage.copy()
    st[0] = 0.0                                  # M = 0 and phi = 0: singular past the wall columns
    bad = sz.BsplineOp.from_storage(bop.k, bop.n, bop.nderiv, bop.kl, bop.ku, st)
    case2 = pc.Case(case.name, bad, case.refs, case.scenario, case.walls, None, case.km, case.kn,
                    case.x, 0j, False)
    got = pc.gpu_invert(case2, "zgbsv", dev)
    assert np.all(got["info"] == pc.oracle_invert(case2, "zgbsv")["info"])
    assert np.allclose(got["x"], case.x.reshape(len(case.km), -1))      # state untouched


def test_invert_window_many_pencils(dev):
    """More pencils than resident CTAs x 2: every slot reuses both of its buffers."""
    case = pc.make_case("tiny_16x24x16", npencils=6000)
    got = pc.gpu_invert(case, "zgbsv", dev)
    want = pc.oracle_invert(case, "zgbsv", nthreads=8)
    assert np.all(got["info"] == 0)
    assert np.array_equal(got["ipiv"], want["ipiv"])
    assert pc.relmax(got["x"], want["x"]) <= TOL
    # per-pencil check so that a single bad slot cannot hide behind the global max
    err = np.abs(got["x"] - want["x"]).max(axis=1) / np.abs(want["x"]).max(axis=1)
    assert err.max() <= 1e-11


def test_invert_window_matches_v1(dev, ch96, monkeypatch):
    import subprocess, sys, os, json
    got = pc.gpu_invert(ch96, "zgbsv", dev)
    # the generic kernel is selected per process (SZB_INVERT=v1): run it in a child
    code = (
        "import sys, numpy as np, torch; sys.path.insert(0, %r); import parity_common as pc;"
        "c = pc.make_case('channel_192x96x192', max_pencils=40);"
        "g = pc.gpu_invert(c, 'zgbsv', torch.device('cuda:0'));"
        "np.save(sys.argv[1], g['x']); np.save(sys.argv[2], g['ipiv'])"
    ) % os.path.dirname(os.path.abspath(__file__))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        fx, fp = os.path.join(d, "x.npy"), os.path.join(d, "p.npy")
        env = dict(os.environ, SZB_INVERT="v1")
        subprocess.run([sys.executable, "-c", code, fx, fp], check=True, env=env)
        x1, p1 = np.load(fx), np.load(fp)
    assert np.array_equal(got["ipiv"], p1)
    assert pc.relmax(got["x"], x1) <= TOL


def test_invert_singular_reports_info(dev):
    """A zero reference state with phi*L cancelling M cannot be built easily; instead
    poison one pencil's wavenumbers with NaN-free but singular data: Re -> inf makes
    the viscous terms vanish but M keeps the operator regular, so use phi = 0 and a
    zeroed mass matrix through from_storage."""
    import suzerain_b200 as sz
    case = pc.make_case("tiny_16x24x16", max_pencils=6)
    bop = case.bop
    st = bop.storage.copy()
    st[0] = 0.0                                  # M = 0 => (M + 0 L) singular at column 1
    bad = sz.BsplineOp.from_storage(bop.k, bop.n, bop.nderiv, bop.kl, bop.ku, st)
    case2 = pc.Case(case.name, bad, case.refs, case.scenario, case.walls, None, case.km, case.kn,
                    case.x, 0j, False)
    got = pc.gpu_invert(case2, "zgbsv", dev)
    # wall columns carry the enforcer's unit diagonal; the first interior column is singular
    assert np.all(got["info"] == pc.oracle_invert(case2, "zgbsv")["info"])       # zgbtrf's info
    assert np.allclose(got["x"], case.x.reshape(len(case.km), -1))      # state untouched


# ---------------------------------------------------------------------------
# the reference's own golden vectors replayed on the GPU (tests/golden/*.json)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("cplx", [False, True])
def test_golden_bsmbsm_solve_on_gpu(dev, cplx):
    """tests/test_bsmbsm.cpp:801-966 (solve_real / solve_complex): pack, gbsv, unpermute."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import lib as L
    import test_oracle as to
    from oracle import port
    g, A, papt = to.golden_system(cplx)
    N, KL, KU = A["N"], A["KL"], A["KU"]
    lib = L.load()
    ab = torch.from_numpy(np.where(np.isnan(papt), 0, papt).copy()).to(dev)        # (N, 2KL+KU+1)
    BR = np.array(g["BR"])
    rhs = ((-BR + 1j * BR) if cplx else BR.astype(complex)).reshape(1, N)
    x_in = torch.from_numpy(rhs.copy()).to(dev)
    b = torch.zeros_like(x_in)
    one, zero = (C.c_double * 2)(1.0, 0.0), (C.c_double * 2)(0.0, 0.0)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    L.check("zaPxpby", lib.szb_bsmbsm_zaPxpby_batch(b"N", A["S"], A["n"], one, p(x_in), zero, p(b), 1, s))
    ipiv = torch.zeros(N, dtype=torch.int32, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    ld = 2 * KL + KU + 1
    L.check("zgbtrf", lib.szb_zgbtrf_batch(N, KL, KU, p(ab), ld, N * ld, p(ipiv), p(info), 1, s))
    L.check("zgbtrs", lib.szb_zgbtrs_batch(b"N", N, KL, KU, 1, p(ab), ld, N * ld, p(ipiv), p(b), N, N, 1, s))
    x = torch.zeros_like(b)
    L.check("zaPxpby", lib.szb_bsmbsm_zaPxpby_batch(b"T", A["S"], A["n"], one, p(b), zero, p(x), 1, s))
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    XR = np.array(g["XR"])
    want = (1 + 1j) * XR if cplx else XR
    assert np.abs(x.cpu().numpy().reshape(-1) - want).max() <= 1.8e4 * np.finfo(float).eps * np.abs(XR).max()
    # pivots as LAPACK's (pinned through the oracle port, itself pinned to SciPy's zgbtrf)
    _, ipiv_want, _ = port.zgbtf2(np.where(np.isnan(papt), 0, papt).T.copy(), N, KL, KU)
    assert np.array_equal(ipiv.cpu().numpy(), ipiv_want)


def test_golden_bsplineop_actions_on_gpu(dev):
    """tests/test_bsplineop.cpp:697-742: k = 8 operator actions through the batched apply."""
    import json, os, torch
    import suzerain_b200 as sz
    from suzerain_b200 import lib as L
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bsplineop_colloc.json")))["septic"]
    op = sz.BsplineOp.from_breakpoints(8, np.array(g["breakpoints"]), nderiv=2)
    lib = L.load()
    for act in g["actions"]:
        x = torch.from_numpy(np.array(act["x"]).reshape(act["nrhs"], -1).astype(np.complex128)).to(dev)
        y = torch.zeros_like(x)
        a = (C.c_double * 2)(act["alpha"], 0.0)
        z = (C.c_double * 2)(0.0, 0.0)
        rc = lib.szb_bsplineop_accumulate_complex_batch(op.handle, act["d"], act["nrhs"], a,
                                                        C.c_void_p(x.data_ptr()), op.n, z,
                                                        C.c_void_p(y.data_ptr()), op.n,
                                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
        L.check("bsplineop_accumulate", rc)
        torch.cuda.synchronize()
        want = np.array(act["y"]).reshape(act["nrhs"], -1)
        got = y.cpu().numpy()
        assert np.abs(got.imag).max() == 0
        assert np.abs(got.real - want).max() <= 1000 * np.finfo(float).eps * np.abs(want).max()


# ---------------------------------------------------------------------------
# linearize::rhome_y: the wavenumber-independent "00" operator against the reference's
# own suzerain_rholut_imexop_{accumulate,packf}00 and its rhome_y loop (oracle/_ref)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("casename", ["tiny", "ch96", "bl"])
def test_rhome_y_accumulate_and_invert_match_reference(dev, request, casename):
    import torch
    import suzerain_b200 as sz
    case = request.getfixturevalue(casename)
    P = pc.oracle_problem(case, "ref")
    npen = len(case.km)
    x = case.x.reshape(npen, -1)
    op = pc.make_imexop(case).set_linearization("rhome_y")
    km = torch.from_numpy(case.km).to(dev)                      # ignored under rhome_y
    kn = torch.from_numpy(case.kn).to(dev)
    # accumulate
    beta = 0.3 - 0.8j
    y0 = np.random.default_rng(4).standard_normal(x.shape) + 0j
    want = P.accumulate00(case.phi, x, beta=beta, y=y0)
    xs = torch.from_numpy(case.x).to(dev)
    out = torch.from_numpy(y0.reshape(case.x.shape).copy()).to(dev)
    op.accumulate_batch(case.phi, km, kn, xs, beta, out)
    torch.cuda.synchronize()
    assert pc.relmax(out.cpu().numpy().reshape(npen, -1), want) <= TOL
    # invert, both solver specifications
    want = P.invert00(case.phi, x, want_ipiv=True)
    assert want["info"] == 0
    for solver in ("zgbsv", "zcgbsvx"):
        st = torch.from_numpy(case.x.copy()).to(dev)
        ipiv = torch.zeros((npen, op.N), dtype=torch.int32, device=dev)
        info = torch.full((npen,), -7, dtype=torch.int32, device=dev)
        op.invert_batch(sz.SolverSpec(method=solver), case.phi, km, kn, st, ipiv=ipiv, info=info)
        torch.cuda.synchronize()
        assert np.all(info.cpu().numpy() == 0)
        assert np.array_equal(ipiv.cpu().numpy(), np.tile(want["ipiv"], (npen, 1))), "pivot choices differ"
        assert pc.relmax(st.cpu().numpy().reshape(npen, -1), want["x"]) <= TOL


def test_rhome_y_per_pencil_wrappers_match_reference(dev, tiny):
    """szb_rholut_imexop_{accumulate,packc,packf}00 with the reference's signatures."""
    from suzerain_b200 import lib as L
    case = tiny
    P = pc.oracle_problem(case, "ref")
    lib = L.load()
    op = pc.make_imexop(case)
    n, N = case.n, 5 * case.n
    phi2 = (C.c_double * 2)(case.phi.real, case.phi.imag)
    scen = L.Scenario(*[case.scenario[k] for k in ("Re", "Pr", "Ma", "alpha", "gamma")])
    refs = np.ascontiguousarray(case.refs)
    ref = L.Ref(*[refs[i].ctypes.data_as(L.c_double_p) for i in range(26)])
    refld = L.RefLd(*([1] * 26))
    A = lib.szb_imexop_bsmbsm(op.handle)
    for packf, fn in ((False, lib.szb_rholut_imexop_packc00), (True, lib.szb_rholut_imexop_packf00)):
        rows = A.LD + (A.KL if packf else 0)
        out = np.full((N, rows), np.nan + 1j * np.nan, dtype=np.complex128)
        L.check("pack00", fn(phi2, C.byref(scen), C.byref(ref), C.byref(refld), case.bop.handle, C.byref(A),
                             out.ctypes.data_as(C.c_void_p), None))
        want = P.assemble00(case.phi, packf=packf, with_bc=False)
        m = ~np.isnan(want)
        assert np.array_equal(np.isnan(out), np.isnan(want))
        assert np.abs(out[m] - want[m]).max() <= TOL * np.abs(want[m]).max()
    x = case.x[0]
    y = np.zeros_like(x)
    b2 = (C.c_double * 2)(0.0, 0.0)
    xin = [np.ascontiguousarray(x[f]) for f in range(5)]
    yout = [np.ascontiguousarray(y[f]) for f in range(5)]
    L.check("accumulate00", lib.szb_rholut_imexop_accumulate00(
        phi2, C.byref(scen), C.byref(ref), C.byref(refld), case.bop.handle,
        *[a.ctypes.data_as(C.c_void_p) for a in xin], b2, *[a.ctypes.data_as(C.c_void_p) for a in yout], None))
    want = P.accumulate00(case.phi, x.reshape(1, -1))
    assert pc.relmax(np.concatenate(yout).reshape(1, -1), want) <= TOL


# ---------------------------------------------------------------------------
# wave-space building blocks of the nonlinear operator: batched B-spline operator
# apply and diffwave against the reference's own C (oracle/_ref)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("k,Ny,nrhs", [(8, 96, 1000), (6, 40, 37), (4, 24, 5), (10, 64, 129)])
@pytest.mark.parametrize("beta", [0.0, 0.7 - 0.2j])
def test_bsplineop_accumulate_batch_matches_reference(dev, k, Ny, nrhs, beta):
    import torch
    import suzerain_b200 as sz
    from oracle import ref as oref
    case = pc.make_case("tiny_16x24x16", max_pencils=4, k=k, Ny=Ny)
    P = pc.oracle_problem(case, "ref")
    rng = np.random.default_rng(5)
    x = rng.standard_normal((nrhs, Ny)) + 1j * rng.standard_normal((nrhs, Ny))
    y0 = rng.standard_normal((nrhs, Ny)) + 1j * rng.standard_normal((nrhs, Ny))
    alpha = 1.3 + 0.4j
    for d in range(3):
        want = P.bsplineop_accumulate_complex(d, alpha, x, beta, y0)
        got = sz.bsplineop_accumulate_complex_batch(case.bop, d, alpha, torch.from_numpy(x).to(dev), beta,
                                                    torch.from_numpy(y0.copy()).to(dev))
        torch.cuda.synchronize()
        assert pc.relmax(got.cpu().numpy(), want) <= TOL


@pytest.mark.parametrize("dxcnt,dzcnt", [(0, 0), (1, 0), (0, 1), (2, 0), (0, 2), (1, 1), (2, 1)])
@pytest.mark.parametrize("sub", [False, True])
def test_diffwave_matches_reference(dev, dxcnt, dzcnt, sub):
    """suzerain_diffwave_apply / _accumulate incl. dealiased and Nyquist modes, on the whole wave
    space and on one rank's sub-block of it."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import synth
    from oracle import ref as oref
    Nx, Nz, Ny = 16, 12, 24
    g = sz.wavegrid(Nx, Nz, synth.LX, synth.LZ, xrange=(2, 9) if sub else None, zrange=(5, 17) if sub else None)
    grid = (g.Nx, g.dNx, g.dkbx, g.dkex, g.Nz, g.dNz, g.dkbz, g.dkez)
    nx, nz = g.dkex - g.dkbx, g.dkez - g.dkbz
    rng = np.random.default_rng(9)
    x = rng.standard_normal((nz, nx, Ny)) + 1j * rng.standard_normal((nz, nx, Ny))
    y0 = rng.standard_normal((nz, nx, Ny)) + 1j * rng.standard_normal((nz, nx, Ny))
    alpha, beta = 0.8 - 1.1j, -0.3 + 0.6j
    want = oref.diffwave(dxcnt, dzcnt, alpha, x, synth.LX, synth.LZ, grid)
    got = sz.diffwave_apply(dxcnt, dzcnt, alpha, torch.from_numpy(x.copy()).to(dev), g)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert np.array_equal(got == 0, want == 0), "dealiased / Nyquist pattern differs"
    assert pc.relmax(got, want) <= 4e-16
    want = oref.diffwave(dxcnt, dzcnt, alpha, x, synth.LX, synth.LZ, grid, beta=beta, y=y0)
    got = sz.diffwave_accumulate(dxcnt, dzcnt, alpha, torch.from_numpy(x).to(dev), beta,
                                 torch.from_numpy(y0.copy()).to(dev), g)
    torch.cuda.synchronize()
    assert pc.relmax(got.cpu().numpy(), want) <= 4e-16


# ---------------------------------------------------------------------------
# whole-field HOST-pointer entry points (the three virtuals of operator_hybrid_isothermal)
# ---------------------------------------------------------------------------
def _field_case():
    import suzerain_b200 as sz
    from suzerain_b200 import synth
    case = pc.make_case("tiny_16x24x16")
    g = sz.wavegrid(16, 16, synth.LX, synth.LZ)
    km, kn, act = sz.wavenumbers(g)
    state = synth.state(km, kn, 24, synth.SEED)                       # (npencil, 5, n), dealiased pencils non-zero
    return case, g, km, kn, act, state


@pytest.mark.parametrize("solver", ["zgbsv", "zcgbsvx"])
def test_whole_field_invert_host_api(dev, solver):
    import suzerain_b200 as sz
    case, g, km, kn, act, state = _field_case()
    op = pc.make_imexop(case)
    H = sz.OperatorHybridIsothermal(op, g, sz.SolverSpec(method=solver))
    rng = np.random.default_rng(5)
    N = 5 * case.n
    ic0 = rng.standard_normal((10, N)) + 1j * rng.standard_normal((10, N))     # treatment_constraint.cpp:129-166
    got, got_ic = state.copy(), ic0.copy()
    H.invert_mass_plus_scaled_operator(case.phi, got, ic0=got_ic)
    P = pc.oracle_problem(case)
    flat = state.reshape(len(km), -1)
    want = np.zeros_like(flat)                                          # dealiased pencils zero-filled (:632-637)
    zz = int(np.flatnonzero((km[act] == 0) & (kn[act] == 0))[0])
    extra = np.zeros((act.sum(), 10, N), dtype=complex)
    extra[zz] = ic0
    r = P.invert(solver, case.phi, km[act], kn[act], flat[act], extra=extra)
    want[act] = r["x"]
    assert pc.relmax(got.reshape(len(km), -1), want) <= TOL
    assert np.all(got.reshape(len(km), -1)[~act] == 0)
    assert pc.relmax(got_ic, r["extra"][zz]) <= TOL                    # constraints ride on the (0,0) factorisation


def test_whole_field_apply_and_accumulate_host_api(dev):
    import suzerain_b200 as sz
    case, g, km, kn, act, state = _field_case()
    op = pc.make_imexop(case)
    H = sz.OperatorHybridIsothermal(op, g)
    P = pc.oracle_problem(case)
    flat = state.reshape(len(km), -1)
    # apply: in place, dealiased pencils untouched (operator_hybrid_isothermal.cpp:164-169)
    got = state.copy()
    H.apply_mass_plus_scaled_operator(case.phi, got)
    want = flat.copy()
    want[act] = P.accumulate(case.phi, km[act], kn[act], flat[act])
    assert pc.relmax(got.reshape(len(km), -1), want) <= TOL
    # accumulate: contiguous-state output (field slowest), beta != 0, padded field stride
    npen, n = len(km), case.n
    fs = npen * n + 7
    rng = np.random.default_rng(9)
    out0 = rng.standard_normal(4 * fs + npen * n) + 1j * rng.standard_normal(4 * fs + npen * n)
    out = out0.copy()
    beta = 0.25 + 0.5j
    H.accumulate_mass_plus_scaled_operator(case.phi, state, beta, out, fs)
    y0 = np.stack([out0[f * fs: f * fs + npen * n].reshape(npen, n) for f in range(5)], axis=1).reshape(npen, -1)
    wanty = y0.copy()
    wanty[act] = P.accumulate(case.phi, km[act], kn[act], flat[act], beta=beta, y=y0[act])
    goty = np.stack([out[f * fs: f * fs + npen * n].reshape(npen, n) for f in range(5)], axis=1).reshape(npen, -1)
    assert pc.relmax(goty, wanty) <= TOL
    for f in range(4):                                                  # padding between fields untouched
        assert np.array_equal(out[f * fs + npen * n:(f + 1) * fs], out0[f * fs + npen * n:(f + 1) * fs])


def test_hundred_substeps_track_the_oracle(dev, tiny):
    """north_star: <= 1e-9 relative after 100 time steps.  Drive 100 L-substeps
    (accumulate with the SMR91 alpha, exchange, invert with the SMR91 beta) on the GPU
    and through the oracle from the same state."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import synth
    op = pc.make_imexop(tiny)
    km = torch.from_numpy(tiny.km).to(dev); kn = torch.from_numpy(tiny.kn).to(dev)
    a = torch.from_numpy(tiny.x.copy()).to(dev); tmp = torch.zeros_like(a)
    info = torch.zeros(len(tiny.km), dtype=torch.int32, device=dev)
    P = pc.oracle_problem(tiny)
    ha = tiny.x.reshape(len(tiny.km), -1).copy(); htmp = np.zeros_like(ha)
    dt = synth.delta_t(2.0)
    spec = sz.SolverSpec("zgbsv")
    for step in range(100):
        i = step % 3
        pa, beta, pi = dt * synth.SMR91_ALPHA[i], 1e-3 * dt * synth.SMR91_ZETA[i], -dt * synth.SMR91_BETA[i]
        op.accumulate_batch(pa, km, kn, a, beta, tmp)
        a, tmp = tmp, a
        op.invert_batch(spec, pi, km, kn, a, info=info)
        htmp = P.accumulate(complex(pa), tiny.km, tiny.kn, ha, beta=complex(beta), y=htmp)
        ha, htmp = htmp, ha
        ha = P.invert("zgbsv", complex(pi), tiny.km, tiny.kn, ha, nthreads=4)["x"]
    torch.cuda.synchronize()
    assert int(info.abs().max()) == 0
    assert pc.relmax(a.cpu().numpy().reshape(len(tiny.km), -1), ha) <= 1e-9


# ---------------------------------------------------------------------------
# BASELINE.json's full sizes: oracle on a handful of pencils, size-independent
# properties on everything
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("config,npen", [("channel_1536x384x1152", 6), ("bl_1024x256x512", 6)])
def test_large_ny_matches_oracle(dev, config, npen):
    """N = 1920 (Ny = 384) and N = 1280 with NRBC (Ny = 256): few pencils, full wall-normal size."""
    case = pc.make_case(config, max_pencils=npen)
    for solver in ("zgbsv", "zcgbsvx"):
        got = pc.gpu_invert(case, solver, dev)
        want = pc.oracle_invert(case, solver, nthreads=6)
        assert want["info"] == 0 and np.all(got["info"] == 0)
        assert np.array_equal(got["ipiv"], want["ipiv"])
        assert pc.relmax(got["x"], want["x"]) <= TOL
    assert pc.relmax(pc.gpu_accumulate(case, dev), pc.oracle_accumulate(case)) <= TOL


def test_full_grid_properties(dev):
    """Whole wave space of the 192x96x192 channel (18 336 active pencils) through the
    device-resident operator: (i) info == 0 everywhere, (ii) dealiased pencils zero-filled,
    (iii) linearity of the solve, (iv) accumulate o invert = identity off the wall rows,
    (v) a strided sample of pencils against the oracle."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import synth
    Nx, Ny, Nz, k, htdelta, _ = synth.CONFIGS["channel_192x96x192"]
    case = pc.make_case("channel_192x96x192", max_pencils=4)           # operators / profiles / phi
    op = pc.make_imexop(case)
    g = sz.wavegrid(Nx, Nz, synth.LX, synth.LZ)
    H = sz.OperatorHybridIsothermalDevice(op, g, sz.SolverSpec("zgbsv"), dev)
    assert H.nactive == 96 * 191 and H.npencil == 145 * 288
    rng = np.random.default_rng(17)
    shape = (H.npencil, 5, Ny)
    x = torch.from_numpy(rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).to(dev)
    y = torch.from_numpy(rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).to(dev)
    a, b = 0.7 - 0.2j, -1.3 + 0.4j
    phi = case.phi
    sx, sy, sc = x.clone(), y.clone(), (a * x + b * y)
    ipiv = torch.zeros((H.nactive, 5 * Ny), dtype=torch.int32, device=dev)
    H.invert_mass_plus_scaled_operator(phi, sx, ipiv=ipiv)
    assert int(H.info.abs().max()) == 0
    H.invert_mass_plus_scaled_operator(phi, sy)
    H.invert_mass_plus_scaled_operator(phi, sc)
    torch.cuda.synchronize()
    inact = torch.from_numpy(H.h_inactive.astype(np.int64)).to(dev)
    act = torch.from_numpy(H.h_active.astype(np.int64)).to(dev)
    assert float(sx[inact].abs().max()) == 0.0                           # (ii)
    lin = (sc[act] - (a * sx[act] + b * sy[act])).abs().max() / sc[act].abs().max()
    assert float(lin) <= 1e-12                                           # (iii)
    back = torch.zeros_like(x)
    H.op.accumulate_batch(phi, H.km, H.kn, sx, 0.0, back, index=H.active)
    torch.cuda.synchronize()
    d = (back[act] - x[act]).abs()
    scale = float(x.abs().max())
    assert float(d[:, :, 1:-1].max()) / scale <= 1e-11                   # (iv) interior points
    assert float(d[:, 4, :].max()) / scale <= 1e-11                      #      and the continuity rows at the walls
    sel = np.linspace(0, H.nactive - 1, 24).astype(int)                 # (v)
    P = pc.oracle_problem(case)
    hx = x.cpu().numpy().reshape(H.npencil, -1)[H.h_active[sel]]
    want = P.invert("zgbsv", phi, H.h_km[sel], H.h_kn[sel], hx, want_ipiv=True, nthreads=8)
    got = sx.cpu().numpy().reshape(H.npencil, -1)[H.h_active[sel]]
    assert pc.relmax(got, want["x"]) <= TOL
    assert np.array_equal(ipiv.cpu().numpy()[sel], want["ipiv"])


# ---------------------------------------------------------------------------
# The remaining solver specifications of apps/perfect/test_implicit_solvers.sh:24-30
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("casename,scale,equil", [("tiny", 1, False), ("tiny", 1, True), ("tiny", 100, True),
                                                  ("tiny", 1000, True), ("ch96", 1, False), ("ch96", 1000, True),
                                                  ("bl", 1, False), ("bl", 100, True)])
def test_zgbsvx_matches_lapack_expert_driver(dev, request, casename, scale, equil):
    """zgbsvx[,equil=b]: LAPACK's expert driver as the reference calls it (bsmbsm_solver.cpp:253-289).
    Scaling phi by 100 / 1000 makes zlaqgb choose row / row-and-column equilibration."""
    import dataclasses
    import suzerain_b200 as sz
    base = request.getfixturevalue(casename)
    case = dataclasses.replace(base, phi=base.phi * scale)
    npen = len(case.km)
    P = pc.oracle_problem(case, "ref")
    want = P.invert_spec(case.phi, case.km, case.kn, case.x.reshape(npen, -1), method="zgbsvx", equil=equil)
    assert want["info"] == 0
    equed = set(want["stats"][:, 0].astype(int).tolist())
    if casename == "tiny":
        assert equed == ({0} if scale == 1 else {1} if scale == 100 else {3}), equed
    got = pc.gpu_invert(case, "zgbsvx", dev, spec=sz.SolverSpec(method="zgbsvx", equil=equil))
    assert np.all(got["info"] == 0)
    assert pc.relmax(got["x"], want["x"]) <= TOL
    # and the plain zgbsv pivots when nothing was scaled (same factorisation)
    if equed == {0}:
        ref = pc.oracle_invert(case, "zgbsv")
        assert np.array_equal(got["ipiv"], ref["ipiv"])


@pytest.mark.parametrize("text", ["zcgbsvx,reuse=true,aiter=1,siter=-1,diter=5,tolsc=0",
                                  "zcgbsvx,reuse=true,aiter=5,siter=25,diter=5,tolsc=0"])
def test_reuse_and_siter_hints_stay_within_the_reference_chain(dev, text):
    """reuse=true / siter>=0 are hints on the device.  The reference, carrying its factorisation
    along a kx row as a preconditioner (and trying single precision first), refines to the same
    stopping criterion; its own test holds all variants to 6e-15 of each other in the restart
    fields.  Here: two full kx rows of the channel grid, reference chain vs device."""
    import dataclasses
    import suzerain_b200 as sz
    full = pc.make_case("channel_192x96x192")
    rowlen = 96
    assert np.ptp(full.kn[:rowlen]) == 0.0 and full.kn[rowlen] != full.kn[0]      # kx inner, kz outer
    case = dataclasses.replace(full, km=full.km[:2 * rowlen], kn=full.kn[:2 * rowlen], x=full.x[:2 * rowlen])
    spec = sz.SolverSpec.parse(text)
    P = pc.oracle_problem(case, "ref")
    x = case.x.reshape(2 * rowlen, -1)
    want = P.invert_spec(case.phi, case.km, case.kn, x, method="zcgbsvx", reuse=spec.reuse, aiter=spec.aiter,
                         siter=spec.siter, diter=spec.diter, tolsc=spec.tolsc, rowlen=rowlen, nthreads=2)
    assert want["info"] == 0
    assert want["stats"][:, 0].sum() > 0, "the reference never took its approximate-factorisation path"
    got = pc.gpu_invert(case, "zcgbsvx", dev, spec=spec)
    assert np.all(got["info"] == 0)
    assert pc.relmax(got["x"], want["x"]) <= 5e-12
    # per pencil, the device result satisfies the system at least as well as the reference chain's
    fresh = P.invert_spec(case.phi, case.km, case.kn, x, method="zcgbsvx", rowlen=1, nthreads=2)
    assert pc.relmax(got["x"], fresh["x"]) <= TOL


@pytest.mark.parametrize("casename", ["tiny", "ch96"])
def test_rhome_y_refinement_and_extra_right_hand_sides(dev, request, casename):
    """Under linearize::rhome_y every pencil shares one factorisation: zcgbsvx / zgbsvx refine around
    it, and the integral-constraint columns are simply more right hand sides."""
    import torch
    import suzerain_b200 as sz
    case = request.getfixturevalue(casename)
    P = pc.oracle_problem(case, "ref")
    npen = len(case.km)
    x = case.x.reshape(npen, -1)
    want = P.invert00(case.phi, x)
    assert want["info"] == 0
    op = pc.make_imexop(case).set_linearization("rhome_y")
    km = torch.from_numpy(case.km).to(dev)
    kn = torch.from_numpy(case.kn).to(dev)
    for text in ("zcgbsvx", "zgbsvx,equil=false", "zcgbsvx,aiter=2,diter=3"):
        st = torch.from_numpy(case.x.copy()).to(dev)
        info = torch.full((npen,), -7, dtype=torch.int32, device=dev)
        iters = torch.full((npen,), -7, dtype=torch.int32, device=dev)
        op.invert_batch(sz.SolverSpec.parse(text), case.phi, km, kn, st, info=info, iters=iters)
        torch.cuda.synchronize()
        assert np.all(info.cpu().numpy() == 0), text
        it = iters.cpu().numpy()
        assert it.min() >= 0 and it.max() <= 5, (text, it.min(), it.max())
        assert pc.relmax(st.cpu().numpy().reshape(npen, -1), want["x"]) <= TOL, text
    # extra right hand sides with zgbsv
    rng = np.random.default_rng(9)
    nextra = 3
    extra = rng.standard_normal((npen, nextra, 5 * case.n)) + 1j * rng.standard_normal((npen, nextra, 5 * case.n))
    wex = P.invert00(case.phi, extra.reshape(npen * nextra, -1))["x"].reshape(extra.shape)
    st = torch.from_numpy(case.x.copy()).to(dev)
    ex = torch.from_numpy(extra.copy()).to(dev)
    info = torch.full((npen,), -7, dtype=torch.int32, device=dev)
    op.invert_batch(sz.SolverSpec(method="zgbsv"), case.phi, km, kn, st, extra=ex, info=info)
    torch.cuda.synchronize()
    assert np.all(info.cpu().numpy() == 0)
    assert pc.relmax(st.cpu().numpy().reshape(npen, -1), want["x"]) <= TOL
    assert pc.relmax(ex.cpu().numpy(), wex) <= TOL


def test_rhome_y_warp_per_pencil_fallback_matches_reference(dev):
    """The warp-per-pencil sweeps (shapes the thread-per-pencil kernel has no instantiation for) are
    selected by SZB_RHOME_Y_WARP, read once per process: run them in a child process."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "tests")
import parity_common as pc, suzerain_b200 as sz
case = pc.make_case("tiny_16x24x16")
P = pc.oracle_problem(case, "ref")
npen = len(case.km)
want = P.invert00(case.phi, case.x.reshape(npen, -1), want_ipiv=True)
op = pc.make_imexop(case).set_linearization("rhome_y")
dev = torch.device("cuda:0")
st = torch.from_numpy(case.x.copy()).to(dev)
km = torch.from_numpy(case.km).to(dev); kn = torch.from_numpy(case.kn).to(dev)
ipiv = torch.zeros((npen, op.N), dtype=torch.int32, device=dev)
info = torch.full((npen,), -7, dtype=torch.int32, device=dev)
op.invert_batch(sz.SolverSpec(method="zgbsv"), case.phi, km, kn, st, ipiv=ipiv, info=info)
torch.cuda.synchronize()
assert np.all(info.cpu().numpy() == 0)
assert np.array_equal(ipiv.cpu().numpy(), np.tile(want["ipiv"], (npen, 1)))
err = pc.relmax(st.cpu().numpy().reshape(npen, -1), want["x"])
assert err <= 1e-12, err
print("ok", err)
'''
    import os
    env = dict(os.environ, SZB_RHOME_Y_WARP="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=pc.ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name", ["channel_k08", "coleman3k01.00"])
def test_reference_restart_state_and_profiles(dev, name):
    """Grid, scenario, mean profiles and mean state of the reference's own restart files
    (fields/channel_k08.h5: Ny=96, k=8, htdelta=3 -- the bench grid; fields/coleman3k01.00.h5: Ny=128)."""
    case = pc.make_case_from_restart(name)
    got = pc.gpu_accumulate(case, dev)
    assert pc.relmax(got, pc.oracle_accumulate(case, kind="ref")) <= TOL
    for solver in ("zgbsv", "zcgbsvx"):
        got = pc.gpu_invert(case, solver, dev)
        want = pc.oracle_invert(case, solver, kind="ref")
        assert want["info"] == 0 and np.all(got["info"] == 0)
        assert np.array_equal(got["ipiv"], want["ipiv"]), "pivot choices differ"
        assert pc.relmax(got["x"], want["x"]) <= TOL


@pytest.mark.parametrize("solver", ["zgbsv", "zcgbsvx"])
def test_nan_right_hand_side_stays_in_its_pencil(dev, tiny, solver):
    """SURVEY 8g-7 (dsgbsvx.def:306-310): a NaN in b makes that pencil's x all NaN; the reference reports
    no error for it.  The other pencils of the batch must not notice."""
    import dataclasses
    x = tiny.x.copy()
    bad = 3
    x[bad, 2, 5] = np.nan
    case = dataclasses.replace(tiny, x=x)
    got = pc.gpu_invert(case, solver, dev)
    want = pc.oracle_invert(case, solver, kind="ref")
    assert np.all(got["info"] == 0) and want["info"] == 0
    assert np.all(np.isnan(want["x"][bad])) and np.all(np.isnan(got["x"][bad].real) | np.isnan(got["x"][bad].imag))
    ok = np.arange(len(case.km)) != bad
    assert not np.isnan(got["x"][ok]).any()
    assert pc.relmax(got["x"][ok], want["x"][ok]) <= TOL
