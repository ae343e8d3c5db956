"""lowstorage::{method, substep, step} mirror (suzerain/lowstorage.hpp:930-1520)."""
import os
import sys
from fractions import Fraction as F

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import parity_common as pc                                    # noqa: E402
from suzerain_b200 import lowstorage as ls, synth             # noqa: E402


def test_smr91_and_yang11_constants():
    m = ls.SMR91
    assert m.substeps == 3 and m.name == "smr91"
    assert [m.fractions("alpha", i) for i in range(3)] == [F(29, 96), F(-3, 40), F(1, 6)]
    assert [m.fractions("beta", i) for i in range(3)] == [F(37, 160), F(5, 24), F(1, 6)]
    assert [m.fractions("gamma", i) for i in range(3)] == [F(8, 15), F(5, 12), F(3, 4)]
    assert [m.fractions("zeta", i) for i in range(3)] == [F(0), F(-17, 60), F(-5, 12)]
    assert [m.eta(i) for i in range(4)] == [0.0, 256 / 480, 320 / 480, 1.0]          # ends at t + dt
    for i in range(3):
        assert abs(m.iota_alpha(i) + m.iota_beta(i) - m.iota(i)) <= 1e-16
        assert (m.alpha(i), m.beta(i), m.zeta(i)) == (synth.SMR91_ALPHA[i], synth.SMR91_BETA[i], synth.SMR91_ZETA[i])
    assert m.evmaxmag_imag() == np.sqrt(3.0) and abs(m.evmaxmag_real() - 2.512745326618329) < 1e-15
    y = ls.YANG11
    assert [y.fractions("alpha", i) for i in range(3)] == [F(1, 3), F(-1, 2), F(1, 3)]
    assert [y.fractions("beta", i) for i in range(3)] == [F(1, 6), F(2, 3), F(0)]
    assert y.eta(3) == 1.0
    assert ls.delta_t_reducer([0.3, 0.1, 0.2]) == 0.1 and np.isnan(ls.delta_t_reducer([0.3, float("nan")]))


class OracleL:
    """The three virtuals on numpy states through the oracle (checker side of the tests)."""

    def __init__(self, case):
        self.case, self.P = case, pc.oracle_problem(case, "ref")
        self.calls = []

    def apply_mass_plus_scaled_operator(self, phi, state):
        self.calls.append("apply")
        state[...] = self.P.accumulate(complex(phi), self.case.km, self.case.kn, state.reshape(len(state), -1)).reshape(state.shape)

    def accumulate_mass_plus_scaled_operator(self, phi, input, beta, output):
        self.calls.append("accumulate")
        output[...] = self.P.accumulate(complex(phi), self.case.km, self.case.kn, input.reshape(len(input), -1),
                                        beta=complex(beta), y=output.reshape(len(output), -1)).reshape(output.shape)

    def invert_mass_plus_scaled_operator(self, phi, state):
        self.calls.append("invert")
        r = self.P.invert("zgbsv", complex(phi), self.case.km, self.case.kn, state.reshape(len(state), -1), nthreads=4)
        assert r["info"] == 0
        state[...] = r["x"].reshape(state.shape)


class ToyN:
    """A stand-in for the nonlinear operator: state <- w * state with a fixed real weight field
    (keeps the mean mode real), stable-step candidates from a fixed list."""

    def __init__(self, shape, xp, dt=2e-3):
        rng = np.random.default_rng(12)
        self.w = 0.5 + rng.random(shape)
        self.xp, self.dt, self.times = xp, dt, []

    def apply_operator(self, time, state, method, substep_index):
        self.times.append(time)
        if self.xp is np:
            state.data *= self.w
        else:
            import torch
            if not hasattr(self, "wd"):
                self.wd = torch.from_numpy(self.w).to(state.data.device)
            state.data.mul_(self.wd)
        return [3.0 * self.dt, self.dt, 2.0 * self.dt]


def test_step_is_one_apply_two_accumulate_three_invert_and_matches_the_written_out_scheme():
    pytest.importorskip("scipy")
    case = pc.make_case("tiny_16x24x16", max_pencils=6)
    m, chi = ls.SMR91, 0.37
    L, N = OracleL(case), ToyN(case.x.shape, np)
    a, b = ls.State(case.x.copy()), ls.State(np.zeros_like(case.x))
    dt = ls.step(m, ls.delta_t_reducer, L, chi, N, 1.5, a, b)
    assert dt == N.dt
    assert L.calls == ["apply", "invert", "accumulate", "invert", "accumulate", "invert"]
    assert np.allclose(N.times, [1.5 + dt * m.eta(i) for i in range(3)], rtol=0, atol=1e-15)
    # the scheme written out by hand:  (M - dt beta_i L) u_{i+1} = (M + dt alpha_i L) u_i + chi dt (gamma_i N(u_i) + zeta_i N(u_{i-1}))
    P, km, kn = L.P, case.km, case.kn
    flat = lambda v: v.reshape(len(v), -1)
    u, Nprev = case.x.copy(), None
    for i in range(3):
        Nu = u * N.w
        rhs = P.accumulate(complex(dt * m.alpha(i)), km, kn, flat(u)).reshape(u.shape) + chi * dt * m.gamma(i) * Nu
        if i > 0:
            rhs += chi * dt * m.zeta(i) * Nprev
        u = P.invert("zgbsv", complex(-dt * m.beta(i)), km, kn, flat(rhs))["x"].reshape(u.shape)
        Nprev = Nu
    assert pc.relmax(a.data, u) <= 1e-12
    # substep() reproduces the first substep from the same start (b holds N(u) history = 0 at substep 0)
    a2, b2 = ls.State(case.x.copy()), ls.State(np.zeros_like(case.x))
    L2, N2 = OracleL(case), ToyN(case.x.shape, np)
    ls.substep(m, L2, chi, N2, 1.5, a2, b2, dt, 0)
    assert L2.calls == ["accumulate", "invert"]
    u0 = case.x.copy()
    rhs = P.accumulate(complex(dt * m.alpha(0)), km, kn, flat(u0)).reshape(u0.shape) + chi * dt * m.gamma(0) * (u0 * N2.w)
    want = P.invert("zgbsv", complex(-dt * m.beta(0)), km, kn, flat(rhs))["x"].reshape(u0.shape)
    assert pc.relmax(b2.data, want) <= 1e-12
    with pytest.raises(ValueError):
        ls.substep(m, L2, chi, N2, 0.0, a2, b2, dt, 3)


@pytest.mark.gpu
def test_gpu_steps_track_the_oracle_through_the_same_driver():
    """Four full SMR91 steps (12 substeps) of the device-resident operator against the oracle, both
    driven by lowstorage.step with the same stand-in nonlinear operator."""
    import torch
    import suzerain_b200 as sz
    dev = torch.device("cuda:0")
    Nx, Ny, Nz, k, htdelta, _ = synth.CONFIGS["tiny_16x24x16"]
    base = pc.make_case("tiny_16x24x16")
    g = sz.wavegrid(Nx, Nz, synth.LX, synth.LZ)
    km, kn, act = sz.wavenumbers(g)
    x_all = synth.state(km, kn, Ny, synth.SEED)
    x_all[~act] = 0
    import dataclasses
    case = dataclasses.replace(base, km=km[act], kn=kn[act], x=x_all[act])
    m, chi = ls.SMR91, 1.0 / (g.dNx * g.dNz)
    # oracle side (active pencils only: dealiased ones stay zero on both sides)
    Lo, No = OracleL(case), ToyN(case.x.shape, np, dt=1e-3)
    ao, bo = ls.State(case.x.copy()), ls.State(np.zeros_like(case.x))
    # device side: the whole wave space, interleaved buffers
    H = sz.OperatorHybridIsothermalDevice(pc.make_imexop(case), g, sz.SolverSpec(method="zgbsv"), dev)
    Ld = ls.InterleavedOperator(H)
    Nd = ToyN(case.x.shape, torch, dt=1e-3)
    w_all = np.ones(x_all.shape); w_all[act] = Nd.w
    Nd.w = w_all
    ad, bd = ls.State(torch.from_numpy(x_all.copy()).to(dev)), ls.State(torch.zeros(x_all.shape, dtype=torch.complex128, device=dev))
    t = 0.0
    for _ in range(4):
        dto = ls.step(m, ls.delta_t_reducer, Lo, chi, No, t, ao, bo)
        dtd = ls.step(m, ls.delta_t_reducer, Ld, chi, Nd, t, ad, bd)
        assert dto == dtd
        t += dto
    torch.cuda.synchronize()
    got = ad.data.cpu().numpy()
    assert int(H.info.abs().max()) == 0
    assert np.all(got[~act] == 0)
    assert pc.relmax(got[act], ao.data) <= 1e-11


def test_step_honours_max_delta_t_and_propagates_nan():
    """math::minnan semantics of lowstorage::step (lowstorage.hpp:1495-1499) with do-nothing operators."""
    class L0:
        def __init__(self):
            self.phis = []

        def apply_mass_plus_scaled_operator(self, phi, state):
            self.phis.append(("apply", phi))

        def accumulate_mass_plus_scaled_operator(self, phi, input, beta, output):
            self.phis.append(("accumulate", phi, beta))

        def invert_mass_plus_scaled_operator(self, phi, state):
            self.phis.append(("invert", phi))

    class N0:
        def __init__(self, cands):
            self.cands = cands

        def apply_operator(self, time, state, method, substep_index):
            return self.cands

    m = ls.SMR91
    a, b = ls.State(np.ones(3)), ls.State(np.zeros(3))
    L = L0()
    dt = ls.step(m, ls.delta_t_reducer, L, 0.5, N0([0.4, 0.2]), 0.0, a, b, max_delta_t=0.05)
    assert dt == 0.05
    assert L.phis[0] == ("apply", 0.05 * m.alpha(0)) and L.phis[1] == ("invert", -0.05 * m.beta(0))
    assert L.phis[2] == ("accumulate", 0.05 * m.alpha(1), 0.5 * 0.05 * m.zeta(1))
    assert L.phis[-1] == ("invert", -0.05 * m.beta(2))
    assert ls.step(m, ls.delta_t_reducer, L0(), 0.5, N0([0.4, 0.2]), 0.0, a, b) == 0.2
    assert np.isnan(ls.step(m, ls.delta_t_reducer, L0(), 0.5, N0([float("nan"), 0.2]), 0.0, a, b, max_delta_t=0.05))
