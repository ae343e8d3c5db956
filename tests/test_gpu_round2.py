"""Round-2 GPU parity tests: entry points that were exported but unverified (szb_zcgbsvx_batch,
szb_zgbtrs_batch('T'), general alpha/beta of zaPxpby, szb_state_exchange,
szb_imexop_set_refs_device), the (kx,kz)-sharded substep on real device shards against the
single-rank result, larger wavenumber samples of the two big BASELINE grids, and a long run on
the bench grid under both solvers.  Every test calls through the C ABI and compares with oracle/
(the reference's own C built into oracle/_ref, or LAPACK from SciPy's OpenBLAS)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_common as pc                                    # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _p(t):
    return C.c_void_p(t.data_ptr())


def _band_systems(case, nsys, with_bc=True):
    """P (M + phi L)^T P^T of the first nsys pencils in LAPACK band storage (LD rows, packc) from the
    reference's own assembly."""
    mats = [pc.oracle_assemble(case, p, packf=False, with_bc=with_bc) for p in range(nsys)]
    return np.stack(mats)                                      # (nsys, N, LD)


# ---------------------------------------------------------------------------
# bsmbsm_solver protocol on pre-assembled matrices
# ---------------------------------------------------------------------------
def test_zgbtrs_batch_transposed_matches_lapack(dev):
    """szb_zgbtrf_batch + szb_zgbtrs_batch('T') -- the only mode the reference uses
    (bsmbsm_solver.cpp:171-179) -- against the reference's own suzerain_lapack_zgbtrf/zgbtrs('T')."""
    import torch
    from suzerain_b200 import lib as L
    from oracle import ref as oref
    case = pc.make_case("tiny_16x24x16", max_pencils=12)
    nsys = len(case.km)
    papt = _band_systems(case, nsys)
    N, LD = papt.shape[1], papt.shape[2]
    KL = KU = (LD - 1) // 2
    ld = 2 * KL + KU + 1
    ab = np.zeros((nsys, N, ld), dtype=np.complex128)
    ab[:, :, KL:] = papt
    rng = np.random.default_rng(7)
    nrhs = 3
    B = rng.standard_normal((nsys, nrhs, N)) + 1j * rng.standard_normal((nsys, nrhs, N))
    lib = L.load()
    d_ab = torch.from_numpy(ab.copy()).to(dev)
    d_b = torch.from_numpy(B.copy()).to(dev)
    ipiv = torch.zeros((nsys, N), dtype=torch.int32, device=dev)
    info = torch.full((nsys,), -1, dtype=torch.int32, device=dev)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check("zgbtrf", lib.szb_zgbtrf_batch(N, KL, KU, _p(d_ab), ld, N * ld, _p(ipiv), _p(info), nsys, s))
    L.check("zgbtrs", lib.szb_zgbtrs_batch(b"T", N, KL, KU, nrhs, _p(d_ab), ld, N * ld, _p(ipiv), _p(d_b), N,
                                           nrhs * N, nsys, s))
    torch.cuda.synchronize()
    assert np.all(info.cpu().numpy() == 0)
    got = d_b.cpu().numpy()
    for i in range(nsys):
        _, piv, X, rinfo = oref.zgbsv_T(N, KL, KU, ab[i], B[i])
        assert rinfo == 0
        assert pc.relmax(got[i], X) <= TOL
        assert np.array_equal(ipiv[i].cpu().numpy(), piv)


def test_zcgbsvx_batch_matches_reference_lapackext(dev):
    """szb_zcgbsvx_batch against the reference's own suzerain_lapackext_zcgbsvx
    (suzerain/blas_et_al/dsgbsvx.def:71-318) with the default specification (fact = 'N', siter < 0,
    aiter = 1, diter = 5, tolsc = 0), TRANS = 'T' and 'N': solution, iteration count, residual."""
    import torch
    from suzerain_b200 import lib as L
    from oracle import ref as oref
    case = pc.make_case("tiny_16x24x16", max_pencils=10)
    nsys = len(case.km)
    papt = _band_systems(case, nsys)
    N, LD = papt.shape[1], papt.shape[2]
    KL = KU = (LD - 1) // 2
    rng = np.random.default_rng(11)
    B = rng.standard_normal((nsys, N)) + 1j * rng.standard_normal((nsys, N))
    lib = L.load()
    rl = oref.lib()
    f = rl.suzerain_lapackext_zcgbsvx
    f.restype = C.c_int
    for trans in (b"T", b"N"):
        d_ab = torch.from_numpy(papt.copy()).to(dev)
        d_afb = torch.zeros((nsys, N, 2 * KL + KU + 1), dtype=torch.complex128, device=dev)
        d_b = torch.from_numpy(B.copy()).to(dev)
        d_x = torch.zeros_like(d_b)
        ipiv = torch.zeros((nsys, N), dtype=torch.int32, device=dev)
        iters = torch.zeros((nsys,), dtype=torch.int32, device=dev)
        res = torch.zeros((nsys,), dtype=torch.float64, device=dev)
        info = torch.full((nsys,), -1, dtype=torch.int32, device=dev)
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = lib.szb_zcgbsvx_batch(trans, N, KL, KU, 1, 5, C.c_double(0.0), _p(d_ab), N * LD, _p(d_afb),
                                   N * (2 * KL + KU + 1), _p(ipiv), _p(d_b), _p(d_x), _p(iters), _p(res), _p(info),
                                   nsys, s)
        L.check("szb_zcgbsvx_batch", rc)
        torch.cuda.synchronize()
        assert np.all(info.cpu().numpy() == 0)
        got, git, gres = d_x.cpu().numpy(), iters.cpu().numpy(), res.cpu().numpy()
        for i in range(nsys):
            ab = np.ascontiguousarray(papt[i])
            afb = np.zeros((N, 2 * KL + KU + 1), dtype=np.complex128)
            piv = np.zeros(N, dtype=np.int32)
            b = B[i].copy(); x = np.zeros(N, dtype=np.complex128); r = np.zeros(N, dtype=np.complex128)
            fact = C.c_char(b"N"); apprx = C.c_int(0); siter = C.c_int(-1); diter = C.c_int(5)
            tolsc = C.c_double(0.0); afrob = C.c_double(-1.0); rres = C.c_double(0.0)
            pp = lambda a: a.ctypes.data_as(C.c_void_p)
            rc2 = f(C.byref(fact), C.byref(apprx), C.c_int(1), C.c_char(trans), C.c_int(N), C.c_int(KL), C.c_int(KU),
                    pp(ab), C.byref(afrob), pp(afb), pp(piv), pp(b), pp(x), C.byref(siter), C.byref(diter),
                    C.byref(tolsc), pp(r), C.byref(rres))
            assert rc2 == 0
            assert pc.relmax(got[i], x) <= TOL
            assert abs(int(git[i]) - diter.value) <= 1              # stagnation may be seen one step apart
            assert gres[i] <= 10 * max(rres.value, 1e-16 * np.abs(B[i]).max() * np.sqrt(N))
            assert np.array_equal(ipiv[i].cpu().numpy(), piv)


@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (0.5 - 2j, 0.0), (1.0, 1.0), (-1.0, 0.25j), (2.5 + 1j, -0.75 + 0.5j),
                                        (0.0, 1.0)])
@pytest.mark.parametrize("trans", ["N", "T"])
def test_zaPxpby_general_alpha_beta(dev, alpha, beta, trans):
    """suzerain_bsmbsm_zaPxpby (bsmbsm_aPxpby_complex.def:37-336) for general scalars, against the
    permutation written out: y[k] <- alpha x[q(k)] + beta y[k] ('N'), q and qinv swapped for 'T'."""
    import torch
    from suzerain_b200 import lib as L
    S, n, nb = 5, 9, 4
    N = S * n
    rng = np.random.default_rng(3)
    x = rng.standard_normal((nb, N)) + 1j * rng.standard_normal((nb, N))
    y = rng.standard_normal((nb, N)) + 1j * rng.standard_normal((nb, N))
    lib = L.load()
    q = np.array([lib.szb_bsmbsm_q(S, n, i) for i in range(N)])
    qinv = np.array([lib.szb_bsmbsm_qinv(S, n, i) for i in range(N)])
    perm = q if trans == "N" else qinv
    want = alpha * x[:, perm] + beta * y
    dx, dy = torch.from_numpy(x.copy()).to(dev), torch.from_numpy(y.copy()).to(dev)
    d2 = lambda z: (C.c_double * 2)(complex(z).real, complex(z).imag)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check("zaPxpby", lib.szb_bsmbsm_zaPxpby_batch(trans.encode(), S, n, d2(alpha), _p(dx), d2(beta), _p(dy), nb, s))
    torch.cuda.synchronize()
    assert np.abs(dy.cpu().numpy() - want).max() <= 4 * np.finfo(float).eps * max(np.abs(want).max(), 1.0)


# ---------------------------------------------------------------------------
# state exchange and device-side reference profiles
# ---------------------------------------------------------------------------
def test_state_exchange_swaps_layouts_bit_exactly(dev):
    """b.exchange(a) (suzerain/lowstorage.hpp:1511, suzerain/state.hpp:486-520,607-630) between the
    interleaved and the contiguous layout, padded contiguous field stride, every stored pencil."""
    import torch
    from suzerain_b200 import lib as L
    npen, n, pad = 37, 24, 5
    rng = np.random.default_rng(5)
    a = rng.standard_normal((npen, 5, n)) + 1j * rng.standard_normal((npen, 5, n))
    b = rng.standard_normal((5, npen + pad, n)) + 1j * rng.standard_normal((5, npen + pad, n))
    da, db = torch.from_numpy(a.copy()).to(dev), torch.from_numpy(b.copy()).to(dev)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check("szb_state_exchange", L.load().szb_state_exchange(npen, None, 5, n, _p(da), n, 5 * n, _p(db),
                                                              (npen + pad) * n, n, s))
    torch.cuda.synchronize()
    ga, gb = da.cpu().numpy(), db.cpu().numpy()
    assert np.array_equal(ga, np.transpose(b[:, :npen], (1, 0, 2)))
    assert np.array_equal(gb[:, :npen], np.transpose(a, (1, 0, 2)))
    assert np.array_equal(gb[:, npen:], b[:, npen:])                 # the padding is untouched


def test_set_refs_device_equals_host_setter(dev):
    """szb_imexop_set_refs_device gathers rows q::u .. q::e_deltarho of the reference's 42 x Ny column-major
    `references` block (apps/perfect/references.hpp:82-125, references.cpp:50-108): same operator as the host
    setter, bit for bit."""
    import torch
    case = pc.make_case("tiny_16x24x16", max_pencils=16)
    want = pc.gpu_accumulate(case, dev)
    op = pc.make_imexop(case)
    op.set_refs(np.zeros_like(case.refs))                            # wipe, then restore through the device path
    blk = np.full((case.n, 42), np.nan)
    blk[:, 5:31] = case.refs.T
    op.set_refs_device(torch.from_numpy(np.ascontiguousarray(blk)).to(dev))
    km, kn = torch.from_numpy(case.km).to(dev), torch.from_numpy(case.kn).to(dev)
    x = torch.from_numpy(case.x).to(dev)
    y = torch.zeros_like(x)
    op.accumulate_batch(case.phi, km, kn, x, 0.0, y)
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy().reshape(len(case.km), -1), want)


# ---------------------------------------------------------------------------
# (kx,kz) sharding on real device shards
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("solver", ["zgbsv", "zcgbsvx"])
def test_sharded_substep_matches_single_rank(dev, world, solver):
    """One substep (accumulate -> exchange -> invert, with the per-step reference-profile reduction) on
    `world` shards of the wave space -- shard.shard_wavegrid, each shard with its own operator context,
    device state and stream, on its own GPU when the box has that many -- equals the single-rank result
    bit for bit: wavenumbers are independent and nothing in L communicates."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import shard, synth
    Nx, Ny, Nz, k, htdelta, _ = synth.CONFIGS["tiny_16x24x16"]
    case = pc.make_case("tiny_16x24x16")
    g = sz.wavegrid(Nx, Nz, synth.LX, synth.LZ)
    km, kn, act = sz.wavenumbers(g)
    npen = len(km)
    state = synth.state(km, kn, Ny, synth.SEED)                          # (npen, 5, Ny)
    pa, beta, pi = 2e-3 * synth.SMR91_ALPHA[1], 1e-4, -2e-3 * synth.SMR91_BETA[1]
    spec = sz.SolverSpec(method=solver)
    ngpu = torch.cuda.device_count()

    def run(grid, st, device, stream, nparts):
        with torch.cuda.device(device), torch.cuda.stream(stream):
            op = pc.make_imexop(case)
            # the profiles arrive as the sum of the ranks' contributions (perfect.cpp:1397)
            blk = np.zeros((Ny, 42)); blk[:, 5:31] = case.refs.T
            part = torch.from_numpy(blk / nparts).to(device)
            total = torch.zeros_like(part)
            for _ in range(nparts):
                total += part
            op.set_refs_device(total, stream=stream)
            H = sz.OperatorHybridIsothermalDevice(op, grid, spec, device)
            a = torch.from_numpy(st.copy()).to(device)
            b = torch.full((5, H.npencil, Ny), 0.5 - 0.25j, dtype=torch.complex128, device=device)
            H.accumulate_mass_plus_scaled_operator(pa, a, beta, b, stream=stream)
            H.exchange(a, b, stream=stream)
            H.invert_mass_plus_scaled_operator(pi, a, stream=stream)
            stream.synchronize()
            assert int(H.info.abs().max()) == 0
            return a.cpu().numpy(), b.cpu().numpy()

    nparts = 4                                                       # a power of two: the partial sums are exact
    want_a, want_b = run(g, state, dev, torch.cuda.Stream(dev), nparts)
    nx = g.dkex - g.dkbx
    got_a, got_b = [], []
    for r in range(world):
        mine = shard.shard_wavegrid(g, r, world)
        lo, hi = (mine.dkbz - g.dkbz) * nx, (mine.dkez - g.dkbz) * nx
        d = torch.device("cuda", r % ngpu) if ngpu >= world else dev
        ga, gb = run(mine, state[lo:hi], d, torch.cuda.Stream(d), nparts)
        got_a.append(ga); got_b.append(gb)
    assert np.array_equal(np.concatenate(got_a), want_a)
    assert np.array_equal(np.concatenate(got_b, axis=1), want_b)
    # and the single-rank result is the oracle's
    P = pc.oracle_problem(case)
    flat = state.reshape(npen, -1)
    y = P.accumulate(pa, km[act], kn[act], flat[act], beta=beta, y=np.full_like(flat[act], 0.5 - 0.25j))
    x = P.invert(solver, pi, km[act], kn[act], y)["x"]
    assert pc.relmax(want_a.reshape(npen, -1)[act], x) <= TOL
    assert np.all(want_a.reshape(npen, -1)[~act] == 0)


# ---------------------------------------------------------------------------
# larger samples of the big BASELINE grids; long run on the bench grid
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("config", ["channel_1536x384x1152", "bl_1024x256x512"])
@pytest.mark.parametrize("solver", ["zgbsv", "zcgbsvx"])
def test_big_grids_64_pencils_match_oracle(dev, config, solver):
    """64+ wavenumber pairs of BASELINE configs 3 and 4 incl. the (0,0) mode and the largest |k|
    (make_case always keeps both ends of the active list): solution to 1e-12, pivots identical."""
    case = pc.make_case(config, max_pencils=64)
    assert len(case.km) >= 64 and case.km[0] == 0 and case.kn[0] == 0
    got = pc.gpu_invert(case, solver, dev)
    want = pc.oracle_invert(case, solver, nthreads=8)
    assert np.all(got["info"] == 0) and want["info"] == 0
    if solver == "zgbsv":
        assert np.array_equal(got["ipiv"], want["ipiv"]), "pivot choices differ"
    assert pc.relmax(got["x"], want["x"]) <= TOL


@pytest.mark.parametrize("solver", ["zgbsv", "zcgbsvx"])
def test_hundred_time_steps_on_the_bench_grid(dev, solver):
    """north_star: <= 1e-9 after 100 TIME STEPS (300 substeps) -- channel_192x96x192 (k = 8, Ny = 96), 200+
    pencils, the SMR91 linear substeps (accumulate with beta = chi dt zeta_i, exchange, invert) with the
    scheme's own coefficients, device against oracle substep by substep on the same evolving inputs."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import synth
    case = pc.make_case("channel_192x96x192", max_pencils=200)
    npen, n = len(case.km), case.n
    assert npen >= 200
    P = pc.oracle_problem(case)
    op = pc.make_imexop(case)
    km, kn = torch.from_numpy(case.km).to(dev), torch.from_numpy(case.kn).to(dev)
    spec = sz.SolverSpec(method=solver)
    a = torch.from_numpy(case.x.copy()).to(dev)
    b = torch.zeros_like(a)
    info = torch.zeros(npen, dtype=torch.int32, device=dev)
    ha = case.x.reshape(npen, -1).copy()
    hb = np.zeros_like(ha)
    dt, chi = synth.delta_t(2.0), 1.0 / (288 * 288)
    worst = 0.0
    for it in range(300):
        i = it % 3
        pa, beta, pi = dt * synth.SMR91_ALPHA[i], chi * dt * synth.SMR91_ZETA[i], -dt * synth.SMR91_BETA[i]
        op.accumulate_batch(pa, km, kn, a, beta, b)
        a, b = b, a
        op.invert_batch(spec, pi, km, kn, a, info=info)
        hb = P.accumulate(pa, case.km, case.kn, ha, beta=beta, y=hb, nthreads=8)
        ha, hb = hb, ha
        ha = P.invert(solver, pi, case.km, case.kn, ha, nthreads=8)["x"]
        if it % 30 == 29 or it == 299:
            torch.cuda.synchronize()
            assert int(info.abs().max()) == 0
            worst = max(worst, pc.relmax(a.cpu().numpy().reshape(npen, -1), ha))
    assert worst <= 1e-9, worst


# ---------------------------------------------------------------------------
# the drop-in library: the reference's own per-pencil symbols (include/suzerain_b200_dropin.h)
# ---------------------------------------------------------------------------
class _Cplx(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class _Workspace(C.Structure):
    """suzerain_bsplineop_workspace (suzerain/bsplineop.h:125-180)."""
    _fields_ = [("method", C.c_int), ("k", C.c_int), ("n", C.c_int), ("nderiv", C.c_int),
                ("kl", C.POINTER(C.c_int)), ("ku", C.POINTER(C.c_int)),
                ("max_kl", C.c_int), ("max_ku", C.c_int), ("ld", C.c_int),
                ("D_T", C.POINTER(C.POINTER(C.c_double)))]


def _reference_workspace(bop):
    """A reference-layout workspace over a copy of the operator storage: D_T[d] points max_ku - ku[d]
    doubles into derivative d's block (bsplineop.c:163-187)."""
    from suzerain_b200 import lib as L
    st = np.ascontiguousarray(bop.storage)                      # (nderiv+1, n, ld)
    kl = (C.c_int * len(bop.kl))(*[int(v) for v in bop.kl])
    ku = (C.c_int * len(bop.ku))(*[int(v) for v in bop.ku])
    base = st.ctypes.data
    ptrs = (C.POINTER(C.c_double) * (bop.nderiv + 1))()
    for d in range(bop.nderiv + 1):
        addr = base + 8 * (d * bop.n * bop.ld + int(bop.max_ku - bop.ku[d]))
        ptrs[d] = C.cast(C.c_void_p(addr), C.POINTER(C.c_double))
    w = _Workspace(0, bop.k, bop.n, bop.nderiv, kl, ku, bop.max_kl, bop.max_ku, bop.ld, ptrs)
    return w, (st, kl, ku, ptrs)                                  # keep-alives


def _dropin():
    path = os.path.join(pc.ROOT, "suzerain_b200", "libsuzerain_b200_dropin.so")
    return C.CDLL(path)


def _ref_structs(case, refs):
    from suzerain_b200 import lib as L
    r, ld = L.Ref(), L.RefLd()
    keep = []
    for q, name in enumerate(L.REF_NAMES):
        col = np.ascontiguousarray(refs[q])
        keep.append(col)
        setattr(r, name, col.ctypes.data_as(L.c_double_p))
        setattr(ld, name, 1)
    s = L.Scenario(*[case.scenario[k] for k in ("Re", "Pr", "Ma", "alpha", "gamma")])
    return s, r, ld, keep


@pytest.mark.parametrize("nrbc", [False, True])
def test_dropin_per_pencil_functions_match_the_reference(dev, nrbc):
    """tests/test_rholut_imexop.cpp:86-302 replayed through the reference's own symbols as the drop-in library
    exports them (complex by value, the five ordering integers, buf): k = 6, n = 10, every one of the 26
    reference profiles switched on alone (and all together), packc and packf against the reference's assembly,
    accumulate against the reference's apply, packf == packc shifted by KL rows, NaN-poisoned storage keeps
    its NaNs outside the band."""
    import suzerain_b200 as sz
    from suzerain_b200 import lib as L
    k, n = 6, 10
    bp = np.linspace(0.0, 2.0, n - k + 2) ** 1.3
    bop = sz.BsplineOp.from_breakpoints(k, bp)
    base = pc.make_case("tiny_16x24x16", max_pencils=4, Ny=n, k=k)
    lib = _dropin()
    w, keep_w = _reference_workspace(bop)
    A = L.load().szb_bsmbsm_construct(5, n, bop.max_kl, bop.max_ku)
    N, KL, LD = A.N, A.KL, A.LD
    rng = np.random.default_rng(42)
    abc = [np.asfortranarray(0.3 * rng.standard_normal((5, 5))).reshape(-1, order="F") for _ in range(3)] if nrbc else None
    km, kn, phi = 0.7, -1.3, complex(-0.02, 0.005)
    x = rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))
    y0 = rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))
    beta = complex(0.6, -0.3)
    pd = lambda arr: None if arr is None else arr.ctypes.data_as(L.c_double_p)
    for which in list(range(26)) + [None]:
        refs = np.zeros((26, n))
        if which is None:
            refs = rng.standard_normal((26, n))
        else:
            refs[which] = rng.standard_normal(n)
        case = pc.Case("dropin", bop, refs, base.scenario, base.walls, tuple(abc) if nrbc else None,
                       np.array([km]), np.array([kn]), x[None].copy(), phi, nrbc)
        P = pc.oracle_problem(case)
        s, r, ld, keep = _ref_structs(case, refs)
        a, b, c = (abc if nrbc else (None, None, None))
        # ---- packc / packf ----
        want_c = P.assemble(phi, km, kn, packf=False, with_bc=False)          # (N, LD)
        got_c = np.full((N, LD), np.nan + 1j * np.nan)
        got_f = np.full((N, LD + KL), np.nan + 1j * np.nan)
        for fn, buf_out in ((lib.suzerain_rholut_imexop_packc, got_c), (lib.suzerain_rholut_imexop_packf, got_f)):
            fn.restype = None
            fn(_Cplx(phi.real, phi.imag), C.c_double(km), C.c_double(kn), C.byref(s), C.byref(r), C.byref(ld),
               C.byref(w), 0, 1, 2, 3, 4, None, C.byref(A), buf_out.ctypes.data_as(C.c_void_p), pd(a), pd(b), pd(c))
        inband = ~np.isnan(want_c.real)
        assert np.array_equal(np.isnan(got_c.real), ~inband)                    # NaNs survive outside the band only
        scale = max(np.abs(want_c[inband]).max(), 1e-300)
        assert np.abs(got_c[inband] - want_c[inband]).max() <= 1e-13 * scale
        assert np.all(np.isnan(got_f[:, :KL].real))                             # the factorisation rows stay untouched
        assert np.array_equal(got_f[:, KL:][inband], got_c[inband])             # packf == packc (test :215-230)
        # ---- accumulate ----
        want_y = P.accumulate(phi, np.array([km]), np.array([kn]), x.reshape(1, -1), beta=beta, y=y0.reshape(1, -1).copy())
        yy = [np.ascontiguousarray(y0[f]) for f in range(5)]
        xx = [np.ascontiguousarray(x[f]) for f in range(5)]
        fn = lib.suzerain_rholut_imexop_accumulate
        fn.restype = None
        pv = lambda arr: arr.ctypes.data_as(C.c_void_p)
        fn(_Cplx(phi.real, phi.imag), C.c_double(km), C.c_double(kn), C.byref(s), C.byref(r), C.byref(ld), C.byref(w),
           pv(xx[0]), pv(xx[1]), pv(xx[2]), pv(xx[3]), pv(xx[4]), _Cplx(beta.real, beta.imag),
           pv(yy[0]), pv(yy[1]), pv(yy[2]), pv(yy[3]), pv(yy[4]), pd(a), pd(b), pd(c))
        got_y = np.concatenate(yy)
        assert pc.relmax(got_y, want_y.reshape(-1)) <= TOL


# ---------------------------------------------------------------------------
# the C++ bsmbsm_solver drop-in (include/suzerain_b200_solver.hpp)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["zgbsv", "zcgbsvx"])
def test_cxx_bsmbsm_solver_protocol(dev, method, tmp_path):
    """supply_B -> PAPT -> supplied_PAPT -> solve('T') -> demand_X through the header-only C++ class
    (suzerain/bsmbsm_solver.hpp:70-330 protocol), compiled here with g++, against the reference's zgbtrf +
    zgbtrs('T') with the permutations written out; two right hand sides."""
    import subprocess
    from suzerain_b200 import lib as L
    from oracle import ref as oref
    case = pc.make_case("tiny_16x24x16", max_pencils=3)
    papt = pc.oracle_assemble(case, 1, packf=False, with_bc=True)          # (N, LD)
    N, LD = papt.shape
    KL = KU = (LD - 1) // 2
    n = case.n
    kl = ku = (KL + 1) // 5 - 1
    nrhs = 2
    rng = np.random.default_rng(9)
    b = rng.standard_normal((nrhs, N)) + 1j * rng.standard_normal((nrhs, N))
    src = os.path.join(pc.ROOT, "tests", "cxx", "solver_protocol.cpp")
    exe = str(tmp_path / "solver_protocol")
    libdir = os.path.join(pc.ROOT, "suzerain_b200")
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-I" + os.path.join(pc.ROOT, "include"), src, "-o", exe,
                    "-L" + libdir, "-lsuzerain_b200", "-Wl,-rpath," + libdir], check=True)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([5, n, kl, ku, nrhs], dtype=np.int32).tofile(f)
        b.astype(np.complex128).tofile(f)
        np.where(np.isnan(papt), 0, papt).astype(np.complex128).tofile(f)
    subprocess.run([exe, fin, fout, method], check=True)
    raw = open(fout, "rb").read()
    info = np.frombuffer(raw[:4], dtype=np.int32)[0]
    ipiv = np.frombuffer(raw[4:4 + 4 * N], dtype=np.int32)
    x = np.frombuffer(raw[4 + 4 * N:], dtype=np.complex128).reshape(nrhs, N)
    assert info == 0
    lib = L.load()
    q = np.array([lib.szb_bsmbsm_q(5, n, i) for i in range(N)])
    ab = np.zeros((N, 2 * KL + KU + 1), dtype=np.complex128)
    ab[:, KL:] = np.where(np.isnan(papt), 0, papt)
    _, piv, X, rinfo = oref.zgbsv_T(N, KL, KU, ab, b[:, q])                    # P b
    want = np.empty_like(X)
    want[:, q] = X                                                             # P^T x
    assert rinfo == 0 and np.array_equal(ipiv, piv)
    assert pc.relmax(x, want) <= TOL


# ---------------------------------------------------------------------------
# the rest of the bsplineop apply / accumulate family: real pencils and the in-place forms
# (suzerain/bsplineop.c:222-258, 299-381), against the reference's own dgbmv
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("k,Ny,nrhs", [(8, 96, 1000), (6, 41, 37), (4, 24, 5), (10, 64, 129)])
def test_bsplineop_real_and_in_place_match_reference(dev, k, Ny, nrhs):
    import torch
    import suzerain_b200 as sz
    case = pc.make_case("tiny_16x24x16", max_pencils=4, k=k, Ny=Ny)
    P = pc.oracle_problem(case, "ref")
    rng = np.random.default_rng(11)
    xr = rng.standard_normal((nrhs, Ny)); yr = rng.standard_normal((nrhs, Ny))
    xc = rng.standard_normal((nrhs, Ny)) + 1j * rng.standard_normal((nrhs, Ny))
    for d in range(3):
        for beta in (0.0, -0.6):
            want = P.bsplineop_accumulate(d, 1.3, xr, beta, yr)
            got = sz.bsplineop_accumulate_batch(case.bop, d, 1.3, torch.from_numpy(xr).to(dev), beta,
                                                torch.from_numpy(yr.copy()).to(dev))
            assert pc.relmax(got.cpu().numpy(), want) <= TOL
        # in place, real (odd Ny: pencils that are not 16-byte aligned) and complex
        got = sz.bsplineop_apply_batch(case.bop, d, -0.7, torch.from_numpy(xr.copy()).to(dev))
        assert pc.relmax(got.cpu().numpy(), P.bsplineop_apply(d, -0.7, xr)) <= TOL
        got = sz.bsplineop_apply_batch(case.bop, d, 2.5, torch.from_numpy(xc.copy()).to(dev))
        assert pc.relmax(got.cpu().numpy(), P.bsplineop_apply(d, 2.5, xc)) <= TOL
    # padded leading dimension: only the first Ny entries of a row are touched
    buf = torch.full((nrhs, Ny + 3), 7.0, dtype=torch.float64, device=dev)
    buf[:, :Ny] = torch.from_numpy(xr).to(dev)
    sz.bsplineop_apply_batch(case.bop, 1, 1.0, buf[:, :Ny])
    assert pc.relmax(buf[:, :Ny].cpu().numpy(), P.bsplineop_apply(1, 1.0, xr)) <= TOL
    assert bool((buf[:, Ny:] == 7.0).all())
    # x == y is refused by the accumulating forms, as in the reference (bsplineop.c:244-247)
    t = torch.from_numpy(xr.copy()).to(dev)
    with pytest.raises(Exception):
        sz.bsplineop_accumulate_batch(case.bop, 0, 1.0, t, 0.0, t)


# ---------------------------------------------------------------------------
# collect_references: the physical-space sweep that produces the operator's reference profiles
# (apps/perfect/perfect.cpp:1266-1400)
# ---------------------------------------------------------------------------
def _physical_state(Ny, Nz, Nx, gamma, Ma, seed=21):
    rng = np.random.default_rng(seed)
    y = np.linspace(-1, 1, Ny)[:, None, None]
    rho = 1.0 + 0.2 * rng.uniform(-1, 1, (Ny, Nz, Nx)) + 0.1 * y
    u = 0.6 * (1 - y * y) + 0.2 * rng.standard_normal((Ny, Nz, Nx))
    v = 0.1 * rng.standard_normal((Ny, Nz, Nx))
    w = 0.15 * rng.standard_normal((Ny, Nz, Nx))
    T = 1.0 + 0.3 * rng.uniform(-1, 1, (Ny, Nz, Nx)) + 0.2 * y * y
    p = rho * T / gamma
    e = p / (gamma - 1) + Ma * Ma * rho * (u * u + v * v + w * w) / 2
    return np.stack([e, rho * u, rho * v, rho * w, rho])


@pytest.mark.parametrize("shape,top", [((24, 12, 18), False), ((24, 12, 18), True), ((17, 5, 7), True),
                                       ((96, 96, 144), False)])
def test_collect_references_matches_oracle(dev, shape, top):
    import torch
    import suzerain_b200 as sz
    from oracle import port
    scen = dict(Re=3000.0, Pr=0.7, Ma=1.5, alpha=0.0, gamma=1.4)
    beta = 2.0 / 3.0
    Ny, Nz, Nx = shape
    s = _physical_state(Ny, Nz, Nx, scen["gamma"], scen["Ma"])
    want, mag = port.collect_references(scen["alpha"], beta, scen["gamma"], scen["Ma"], s, top, with_abs=True)
    chi = 1.0 / (Nz * Nx)
    d = torch.from_numpy(s).to(dev)
    got = sz.collect_references(scen, beta, d, top_is_inviscid=top)
    torch.cuda.synchronize()
    got = got.cpu().numpy().T                                           # (42, Ny)
    # tolerance relative to the sum of magnitudes of each (quantity, plane): the sums themselves may cancel
    err = np.abs(got - want * chi) / np.maximum(mag * chi, 1e-300)
    assert err.max() <= TOL, (err.max(), np.unravel_index(err.argmax(), err.shape))
    if top:
        nu_row = port.REFERENCE_QUANTITIES.index("nu")
        assert np.all(got[nu_row:nu_row + 11, -1] == 0.0) and got[port.REFERENCE_QUANTITIES.index("e_deltarho"), -1] == 0.0
    # same bits on a second run (no atomics)
    again = sz.collect_references(scen, beta, d, top_is_inviscid=top).cpu().numpy().T
    assert np.array_equal(again, got)
    # a rank that owns planes [y0, y0 + ny) only: its columns, zeros elsewhere; the pieces add up
    y0, ny = Ny // 3, Ny // 2
    part = sz.collect_references(scen, beta, d[:, y0:y0 + ny].contiguous(), Ny=Ny, y0=y0,
                                 top_is_inviscid=top).cpu().numpy().T
    # (another launch shape, hence another summation tree: equal to rounding, not to the bit)
    assert np.abs(part[:, y0:y0 + ny] - got[:, y0:y0 + ny]).max() <= TOL * np.abs(mag * chi).max()
    assert np.all(part[:, :y0] == 0.0) and np.all(part[:, y0 + ny:] == 0.0)


def test_collect_references_feeds_the_operator(dev):
    """physical state -> collect_references -> set_refs_device -> accumulate equals the host path that sets the 26
    profiles from the oracle's means; and at the full physical extent of the bench grid the linear rows are plain
    sums (size-independent property)."""
    import torch
    import suzerain_b200 as sz
    from oracle import port
    case = pc.make_case("tiny_16x24x16", max_pencils=8)
    Ny = case.bop.n
    scen = dict(case.scenario)
    beta = 2.0 / 3.0
    s = _physical_state(Ny, 12, 18, scen["gamma"], scen["Ma"], seed=4)
    d = torch.from_numpy(s).to(dev)
    refs42 = sz.collect_references(scen, beta, d)
    want = port.collect_references(scen["alpha"], beta, scen["gamma"], scen["Ma"], s) / (12 * 18)
    first = port.REFERENCE_QUANTITIES.index("u")
    x = torch.from_numpy(case.x).to(dev)
    km, kn = torch.from_numpy(case.km).to(dev), torch.from_numpy(case.kn).to(dev)
    outs = []
    for use_device in (True, False):
        op = pc.make_imexop(case)
        if use_device:
            op.set_refs_device(refs42)
        else:
            op.set_refs(np.ascontiguousarray(want[first:first + 26]))
        y = torch.zeros_like(x)
        op.accumulate_batch(case.phi, km, kn, x, 0.0, y)
        torch.cuda.synchronize()
        outs.append(y.cpu().numpy())
    assert pc.relmax(outs[0], outs[1]) <= TOL
    # full size: 96 x 288 x 288 (the dealiased physical extent of channel_192x96x192)
    g = torch.Generator(device=dev); g.manual_seed(3)
    big = torch.rand((5, 96, 288, 288), dtype=torch.float64, device=dev, generator=g) + 1.0
    big[0] += 20.0                                                      # e large enough for p > 0 at Ma = 1.5, |m| <= 2 sqrt 3
    r = sz.collect_references(scen, beta, big)
    torch.cuda.synchronize()
    for name, f in (("rhoE", 0), ("rhou", 1), ("rhov", 2), ("rhow", 3), ("rho", 4)):
        mean = big[f].sum(dim=(1, 2)) / (288 * 288)
        col = r[:, port.REFERENCE_QUANTITIES.index(name)]
        assert float(((col - mean).abs() / mean.abs()).max()) <= TOL
    assert bool(torch.isfinite(r).all())


# ---------------------------------------------------------------------------
# the whole-field HOST entry points on a rank's sub-grid (what bench.py's e2e leg calls under torchrun):
# kz blocks that contain the dealiased gap, end in it, or consist of dealiased rows only
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("solver", ["zgbsv", "zcgbsvx"])
def test_host_invert_on_kz_blocks_matches_whole_field(dev, solver):
    import suzerain_b200 as sz
    from suzerain_b200 import synth
    case = pc.make_case("tiny_16x24x16")
    g = sz.wavegrid(16, 16, synth.LX, synth.LZ)
    km, kn, act = sz.wavenumbers(g)
    state = synth.state(km, kn, 24, synth.SEED)
    op = pc.make_imexop(case)
    whole = state.copy()
    sz.OperatorHybridIsothermal(op, g, sz.SolverSpec(method=solver)).invert_mass_plus_scaled_operator(case.phi, whole)
    nx = g.dkex - g.dkbx
    nz = g.dkez - g.dkbz
    rows = np.flatnonzero(act.reshape(nz, nx).any(axis=1))
    gap = [r for r in range(nz) if r not in rows]
    assert gap, "the dealiased grid has inactive kz rows"
    # blocks: [0, inside the gap), [inside the gap, gap end + 2), [a block of dealiased rows only], [rest]
    cuts = [0, gap[1], gap[-1] + 3, nz]
    blocks = [(cuts[i], cuts[i + 1]) for i in range(3)] + [(gap[0], gap[-1] + 1)]
    for lo, hi in blocks:
        sub = sz.wavegrid(16, 16, synth.LX, synth.LZ, zrange=(g.dkbz + lo, g.dkbz + hi))
        part = state[lo * nx:hi * nx].copy()
        H = sz.OperatorHybridIsothermal(pc.make_imexop(case), sub, sz.SolverSpec(method=solver))
        H.invert_mass_plus_scaled_operator(case.phi, part)
        assert np.array_equal(part, whole[lo * nx:hi * nx]), (lo, hi)
