"""Shared scaffolding for the parity tests: builds one synthetic case and runs
it through (a) the CUDA path via the C ABI and (b) the oracle."""
from __future__ import annotations

import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import suzerain_b200 as sz                                    # noqa: E402
from suzerain_b200 import synth                               # noqa: E402


@dataclasses.dataclass
class Case:
    name: str
    bop: sz.BsplineOp
    refs: np.ndarray            # (26, n)
    scenario: dict
    walls: dict
    nrbc: tuple | None
    km: np.ndarray
    kn: np.ndarray
    x: np.ndarray               # (npencil, 5, n) complex128
    phi: complex
    one_sided: bool

    @property
    def n(self):
        return self.bop.n

    def bc_dict(self):
        """Enforcer data in the oracle's flat form (operator_hybrid_isothermal.cpp:448-461)."""
        g, Ma = self.scenario["gamma"], self.scenario["Ma"]
        lo, up = self.walls["lower"], self.walls["upper"]
        ef = [w[0] / (g * (g - 1)) + Ma * Ma / 2 * (w[1] ** 2 + w[2] ** 2 + w[3] ** 2) for w in (lo, up)]
        return dict(enforce_lower=int(self.walls["enforce_lower"]),
                    enforce_upper=int(self.walls["enforce_upper"]),
                    E_factor=ef, vel_factor=[list(lo[1:]), list(up[1:])])


def make_case(config="tiny_16x24x16", max_pencils=None, phi=None, seed=synth.SEED,
              nrbc=None, scenario=None, walls=None, Ny=None, k=None, npencils=None) -> Case:
    Nx, Ny0, Nz, k0, htdelta, one_sided = synth.CONFIGS[config]
    Ny = Ny or Ny0
    k = k or k0
    Ly = 2.0
    bp = sz.htstretch_breakpoints(Ny, k, 0.0, Ly, htdelta)
    bop = sz.BsplineOp.from_breakpoints(k, bp)
    scenario = dict(scenario or synth.SCENARIO)
    refs = synth.reference_profiles(bop.greville(), Ly, scenario, one_sided)
    g = sz.wavegrid(Nx, Nz, synth.LX, synth.LZ)
    km, kn, act = sz.wavenumbers(g)
    km, kn = km[act], kn[act]
    if max_pencils is not None and len(km) > max_pencils:
        # deterministic spread including the (0,0) mode and the largest wavenumbers
        sel = np.unique(np.concatenate([[0, len(km) - 1],
                                        np.linspace(0, len(km) - 1, max_pencils).astype(int)]))
        km, kn = km[sel], kn[sel]
    if npencils is not None:
        # synthetic wavenumber list of a requested length (persistent-CTA paths)
        rng = np.random.default_rng(seed + 1)
        km = np.concatenate([[0.0], rng.integers(-40, 41, npencils - 1) * (2 * np.pi / synth.LX)])
        kn = np.concatenate([[0.0], rng.integers(-40, 41, npencils - 1) * (2 * np.pi / synth.LZ)])
    x = synth.state(km, kn, Ny, seed)
    if phi is None:
        phi = complex(-synth.delta_t(Ly) * synth.SMR91_BETA[0], 0.0)
    walls = walls or synth.isothermal_walls(one_sided)
    if nrbc is None and one_sided:
        nrbc = synth.nrbc_matrices(seed)
    return Case(config, bop, refs, scenario, walls, nrbc, km, kn, x, phi, one_sided)


def make_case_from_restart(name="channel_k08", Nx=16, Nz=16, seed=synth.SEED) -> Case:
    """A case on the grid, scenario, mean profiles and mean state of one of the reference's own restart
    files (tests/golden/restart_fixtures.npz, extracted by tests/golden/make_restart_golden.py).  Restart
    files hold B-SPLINE COEFFICIENTS (support::save_coefficients; the bar_* samples likewise,
    support.cpp:655-660): the (0,0) pencil is the stored mean state as it is, the reference profiles
    are the stored mean coefficients evaluated at the collocation points (D0 c), and the other pencils
    of an Nx x Nz wave space are synthetic perturbations."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "restart_fixtures.npz"))
    g = lambda key: G[f"{name}/{key}"]
    k, Ny = int(g("k")), int(g("Ny"))
    bop = sz.BsplineOp.from_breakpoints(k, g("breakpoints_y"))
    scenario = dict(Re=float(g("Re")), Pr=float(g("Pr")), Ma=float(g("Ma")), alpha=float(g("alpha")), gamma=float(g("gamma")))
    D0 = bop.dense(0)
    val = lambda c: D0 @ c                                      # coefficients -> collocation-point values
    u = g("bar_u")
    T = val(g("bar_T")[0])
    refs = synth.reference_profiles_from_means(val(g("bar_rho")[0]), val(u[0]), val(u[1]), val(u[2]), T, val(g("bar_mu")[0]),
                                               scenario)
    walls = dict(enforce_lower=True, enforce_upper=True, lower=(float(T[0]), 0.0, 0.0, 0.0), upper=(float(T[-1]), 0.0, 0.0, 0.0))
    wg = sz.wavegrid(Nx, Nz, float(g("Lx")), float(g("Lz")))
    km, kn, act = sz.wavenumbers(wg)
    km, kn = km[act], kn[act]
    x = 1e-2 * synth.state(km, kn, Ny, seed)
    zero = np.flatnonzero((km == 0) & (kn == 0))
    assert len(zero) == 1
    for f, key in enumerate(("rho_E", "rho_u", "rho_v", "rho_w", "rho")):          # ndx::{e, mx, my, mz, rho}
        x[zero[0], f] = g(key).real
    phi = complex(-synth.delta_t(float(g("Ly"))) * synth.SMR91_BETA[0], 0.0)
    return Case("restart:" + name, bop, refs, scenario, walls, None, km, kn, x, phi, False)


# ---------------------------------------------------------------------------
# CUDA path (through the C ABI)
# ---------------------------------------------------------------------------
def make_imexop(case: Case) -> sz.ImexOp:
    op = sz.ImexOp(case.bop)
    op.set_scenario(**case.scenario)
    op.set_refs(case.refs)
    op.set_isothermal(case.walls["enforce_lower"], case.walls["enforce_upper"],
                      case.walls["lower"], case.walls["upper"])
    if case.nrbc is not None:
        op.set_nrbc(*case.nrbc)
    return op


def gpu_invert(case: Case, solver: str, dev, extra=None, spec=None):
    import torch
    op = make_imexop(case)
    km = torch.from_numpy(case.km).to(dev)
    kn = torch.from_numpy(case.kn).to(dev)
    st = torch.from_numpy(case.x.copy()).to(dev)
    npen = len(case.km)
    ipiv = torch.zeros((npen, op.N), dtype=torch.int32, device=dev)
    info = torch.full((npen,), -7, dtype=torch.int32, device=dev)
    iters = torch.zeros((npen,), dtype=torch.int32, device=dev)
    ex = None if extra is None else torch.from_numpy(np.ascontiguousarray(extra)).to(dev)
    spec = spec or sz.SolverSpec(method=solver)
    op.invert_batch(spec, case.phi, km, kn, st, extra=ex, ipiv=ipiv, info=info, iters=iters)
    torch.cuda.synchronize()
    out = dict(x=st.cpu().numpy().reshape(npen, -1), ipiv=ipiv.cpu().numpy(),
               info=info.cpu().numpy(), iters=iters.cpu().numpy())
    if ex is not None:
        out["extra"] = ex.cpu().numpy()
    return out


def gpu_accumulate(case: Case, dev, beta=0.0, y=None, phi=None):
    import torch
    op = make_imexop(case)
    km = torch.from_numpy(case.km).to(dev)
    kn = torch.from_numpy(case.kn).to(dev)
    x = torch.from_numpy(case.x).to(dev)
    out = torch.zeros_like(x) if y is None else torch.from_numpy(np.ascontiguousarray(y)).to(dev)
    op.accumulate_batch(case.phi if phi is None else phi, km, kn, x, beta, out)
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(len(case.km), -1)


def gpu_pack(case: Case, dev, packf=False, with_bc=False, poison=True):
    import torch
    op = make_imexop(case)
    km = torch.from_numpy(case.km).to(dev)
    kn = torch.from_numpy(case.kn).to(dev)
    rows = op.LD + (op.KL if packf else 0)
    out = torch.full((len(case.km), op.N, rows), float("nan"), dtype=torch.complex128, device=dev)
    op.pack_batch(case.phi, km, kn, out, packf=packf, with_bc=with_bc)
    torch.cuda.synchronize()
    return out.cpu().numpy()


# ---------------------------------------------------------------------------
# Oracle (reference build when present, else the C port)
# ---------------------------------------------------------------------------
def oracle_problem(case: Case, kind=None):
    from oracle import ref as oref
    kind = kind or ("ref" if oref.available() else "port")
    if kind == "ref":
        return oref.Problem(case.bop, case.scenario, case.refs, case.bc_dict(), case.nrbc)
    from oracle import port as oport
    return oport.Problem(case.bop, case.scenario, case.refs, case.bc_dict(), case.nrbc)


def oracle_invert(case: Case, solver: str, extra=None, kind=None, nthreads=4):
    P = oracle_problem(case, kind)
    return P.invert(solver, case.phi, case.km, case.kn, case.x.reshape(len(case.km), -1),
                    extra=extra, nthreads=nthreads, want_ipiv=True, want_iters=True)


def oracle_accumulate(case: Case, beta=0.0, y=None, kind=None, phi=None):
    P = oracle_problem(case, kind)
    yy = None if y is None else np.asarray(y).reshape(len(case.km), -1)
    return P.accumulate(case.phi if phi is None else phi, case.km, case.kn,
                        case.x.reshape(len(case.km), -1), beta=beta, y=yy)


def oracle_assemble(case: Case, p: int, packf=False, with_bc=False, kind=None):
    P = oracle_problem(case, kind)
    return P.assemble(case.phi, float(case.km[p]), float(case.kn[p]), packf=packf, with_bc=with_bc)


def relmax(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
