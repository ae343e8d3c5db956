"""Fixtures extracted from the reference's own restart files (fields/*.h5: written by the real
reference build with GSL B-splines through ESIO/HDF5; tests/golden/make_restart_golden.py).

These pin SURVEY 8a rows 15-16 (the B-spline collocation operators, the htstretch grid, the Greville
points) against actual reference OUTPUTS on production-like grids -- orders 6..9, Ny 32..288,
two-sided and one-sided stretching -- rather than against a restatement."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "restart_fixtures.npz"))
NAMES = [str(n) for n in GOLD["names"]]
EPS = np.finfo(float).eps


def gold(name, key):
    return GOLD[f"{name}/{key}"]


@pytest.mark.parametrize("name", NAMES)
def test_product_bspline_operators_match_the_references_stored_operators(name):
    import suzerain_b200 as sz
    k, Ny, ht, Ly = int(gold(name, "k")), int(gold(name, "Ny")), float(gold(name, "htdelta")), float(gold(name, "Ly"))
    bp = gold(name, "breakpoints_y")
    # support::create_bsplines (support.cpp:288-311): the stretched breakpoints themselves
    assert np.abs(sz.htstretch_breakpoints(Ny, k, 0.0, Ly, ht) - bp).max() <= 4 * EPS * Ly
    bop = sz.BsplineOp.from_breakpoints(k, bp)
    assert (bop.n, bop.max_kl, bop.max_ku) == (Ny, int(gold(name, "kl")), int(gold(name, "ku")))
    assert np.abs(bop.greville() - gold(name, "collocation_points_y")).max() <= 4 * EPS * Ly
    assert np.array_equal(bop.knots(), gold(name, "knots"))
    assert np.abs(bop.integration_weights() - gold(name, "integration_weights")).max() <= 4 * EPS * Ly
    for d in range(3):
        want = gold(name, f"Dy{d}T")
        got = np.asarray(bop.storage[d])
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 5e-13 * np.abs(want).max(), (name, d)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_bspline_operators_match_the_references_stored_operators(name):
    pytest.importorskip("scipy")
    from oracle import bspline as obs
    k = int(gold(name, "k"))
    a = obs.make_bsplineop(k, gold(name, "breakpoints_y"))
    assert np.abs(a.knots - gold(name, "knots")).max() == 0.0
    for d in range(3):
        want = gold(name, f"Dy{d}T")
        assert np.abs(a.storage[d] - want).max() <= 5e-13 * np.abs(want).max(), (name, d)


def test_restart_mean_state_is_consistent_with_its_profiles():
    """Restart files hold B-spline COEFFICIENTS (support::save_coefficients; the bar_* samples are
    "(B-spline coefficient, tensor component, sample number)", support.cpp:655-660): for these
    one-dimensional mean states the (0,0) mode of rho is the stored mean-density sample itself, and the
    collocation-point values follow from the mass matrix D0 (a check of the fixture extraction and of the
    operator orientation: density stays positive and bounded, the wall velocity vanishes)."""
    import suzerain_b200 as sz
    for name in [str(n) for n in GOLD["state_names"]]:
        coef = gold(name, "rho")
        assert np.abs(coef.imag).max() == 0.0
        assert np.array_equal(coef.real, gold(name, "bar_rho")[0]), name
        bop = sz.BsplineOp.from_breakpoints(int(gold(name, "k")), gold(name, "breakpoints_y"))
        D0 = bop.dense(0)
        rho = D0 @ coef.real
        assert rho.min() > 0.5 and rho.max() < 2.0
        # clamped B-splines interpolate their end coefficients: wall values are the first / last coefficient
        assert abs(rho[0] - coef.real[0]) <= 1e-14 and abs(rho[-1] - coef.real[-1]) <= 1e-14
        u = (D0 @ gold(name, "rho_u").real) / rho
        assert abs(u[0]) <= 1e-14 and abs(u[-1]) <= 1e-12                     # no slip
        # momentum over density is the stored mean velocity up to the sampling window (bar_* are running means)
        ubar = D0 @ gold(name, "bar_u")[0]
        assert np.abs(u - ubar).max() <= 2e-3 * np.abs(u).max(), name


@pytest.mark.skipif(not os.path.isdir("/root/reference/fields"), reason="reference tree not present")
def test_h5lite_reads_every_reference_restart_file():
    import glob
    from suzerain_b200.h5lite import H5File
    files = sorted(glob.glob("/root/reference/fields/*.h5"))
    assert len(files) >= 20
    for p in files:
        f = H5File(p)
        assert {"Ny", "k", "rho", "rho_E"} <= set(f.keys())
        Ny = int(f["Ny"][0])
        if "breakpoints_y" in f:                                      # absent from the oldest legacy file
            assert f["breakpoints_y"].shape == (Ny - int(f["k"][0]) + 2,)
        assert f["rho"].shape[-2:] == (Ny, 2)                         # complex = double[2] array datatype
    # the committed fixture equals a fresh read
    f = H5File("/root/reference/fields/channel_k08.h5")
    assert np.array_equal(f["Dy1T"], gold("channel_k08", "Dy1T"))
    assert f.attrs("Dy1T")["kl"][0] == 6


@pytest.mark.skipif(not os.path.isdir("/root/reference/fields"), reason="reference tree not present")
def test_restart_loader_on_the_bench_grid_file():
    """suzerain_b200.restart.load on fields/channel_k08.h5 -- the bench grid and scenario."""
    from suzerain_b200 import restart
    r = restart.load("/root/reference/fields/channel_k08.h5")
    assert (r.Nx, r.Ny, r.Nz, r.k, r.htdelta, r.Ly) == (1, 96, 1, 8, 3.0, 2.0)
    assert r.scenario == dict(Re=3000.0, Pr=0.7, Ma=1.5, alpha=0.0, beta=0.7, gamma=1.4)      # = synth.SCENARIO + BETA_VISC
    bop = r.bsplineop()
    for d in (0, 1, 2):
        want = r.operators[d]
        assert np.abs(np.asarray(bop.storage[d]) - want).max() <= 5e-13 * np.abs(want).max()
    st = r.state()
    assert st.shape == (1, 5, 96) and np.array_equal(st[0, 4].real, gold("channel_k08", "rho").real)
    prof = r.mean_profiles(bop)
    assert prof["bar_rho"].shape == (1, 96) and prof["bar_u"].shape == (3, 96)
    assert abs(prof["bar_u"][0, 0]) <= 1e-14 and 0.5 < prof["bar_rho"].min()


def test_restart_state_translates_wavenumbers_into_the_dealiased_grid():
    """A three-dimensional restart field is (Nz, Nx/2+1, Ny), not dealiased and in order in kz
    (support/field.cpp:132-135); Restart.state scatters it into the dealiased (dNz, dNx/2+1) wave space with
    the reference's wavenumber translation (field.cpp:184-207, inorder.c:75-151).  The shipped fixtures are
    all Nx = Nz = 1, so this case is synthetic."""
    import suzerain_b200 as sz
    from suzerain_b200 import restart
    Nx, Ny, Nz = 8, 5, 6
    rng = np.random.default_rng(0)
    fields = {n: rng.standard_normal((Nz, Nx // 2 + 1, Ny)) + 1j * rng.standard_normal((Nz, Nx // 2 + 1, Ny))
              for n in restart.FIELDS}
    R = restart.Restart(path="synthetic", Nx=Nx, Ny=Ny, Nz=Nz, k=4, htdelta=0.0, Lx=4 * np.pi, Ly=2.0, Lz=2 * np.pi,
                        DAFx=1.5, DAFz=1.5, t=0.0, scenario={}, breakpoints_y=None, collocation_points_y=None,
                        operators={}, fields=fields, samples={})
    dNz, nx = R.wave_extents()
    assert (dNz, nx) == (9, 7)
    st = R.state().reshape(dNz, nx, 5, Ny)
    g = sz.wavegrid(Nx, Nz, R.Lx, R.Lz)
    km, kn, act = sz.wavenumbers(g)
    assert st.shape[0] * st.shape[1] == len(km)
    wz = lambda i, N: i if i < N // 2 + 1 else i - N
    seen = np.zeros((dNz, nx), dtype=bool)
    for iz in range(Nz):
        for ix in range(Nx // 2 + 1):
            w = wz(iz, Nz)
            jz = w if w >= 0 else dNz + w
            for f, n in enumerate(restart.FIELDS):
                assert np.array_equal(st[jz, ix, f], fields[n][iz, ix]), (iz, ix, n)
            seen[jz, ix] = True
    assert np.all(st[~seen] == 0)                                  # the dealiasing band is empty
    # every active pencil of the operator's walk carries a file mode; (kx, kz) agree with the wavenumber tables
    flat_seen = seen.reshape(-1)
    assert np.all(flat_seen[act])
    two_pi_Lz = 2 * np.pi / R.Lz
    assert np.isclose(kn.reshape(dNz, nx)[dNz - 1, 0], -two_pi_Lz)     # kz = -1 sits in the last row
    # a finer file than the target keeps only the modes the target can hold
    s, d = restart.wavenumber_translate(8, 4)
    assert s.tolist() == [0, 1, 2, 7] and d.tolist() == [0, 1, 2, 3]
