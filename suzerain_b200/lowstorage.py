"""Low-storage hybrid implicit/explicit time advance: the host-side orchestration that drives the
implicit operator.  Mirrors ``suzerain::lowstorage`` for

    M u_t = L u + chi N(u, t)

* ``Method`` / ``SMR91`` / ``YANG11``   -- ``lowstorage::method<Scheme>`` with the ``smr91`` and ``yang11``
  schemes (suzerain/lowstorage.hpp:930-1230, 1242-1372): alpha, beta, gamma from integer numerators,
  zeta_i = alpha_i + beta_i - gamma_i, eta_i = sum_{j<i} (alpha_j + beta_j), iota, iota_alpha, iota_beta;
* ``substep`` / ``step``                -- lowstorage.hpp:1403-1435 / 1471-1520, same call sequence and
  the same scalar factors: per step one apply + two accumulate + three invert of (M + phi L);
* ``State``                             -- the three state operations the advance needs
  (``assign_from``, ``add_scaled``, ``exchange``; suzerain/state.hpp) on a tensor / array.

``L`` is any object with the three virtuals of ``lowstorage::linear_operator`` (here
``OperatorHybridIsothermalDevice``: device-resident; ``OperatorHybridIsothermal``: host pointers), ``N``
any object with ``apply_operator(time, state, method, substep_index) -> stable time step candidates``.
Nothing here computes: every flop is in the operators.
"""
from __future__ import annotations

import math
from fractions import Fraction


class Method:
    def __init__(self, name, alpha_numerator, beta_numerator, gamma_numerator, denominator,
                 evmaxmag_real, evmaxmag_imag):
        assert len(alpha_numerator) == len(beta_numerator) == len(gamma_numerator)
        self.name = name
        self.substeps = len(alpha_numerator)
        self._a, self._b, self._g, self._den = tuple(alpha_numerator), tuple(beta_numerator), tuple(gamma_numerator), denominator
        self._evr, self._evi = evmaxmag_real, evmaxmag_imag

    def evmaxmag_real(self):
        return self._evr

    def evmaxmag_imag(self):
        return self._evi

    def _eta_numerator(self, i):
        assert 0 <= i <= self.substeps                       # i == substeps is allowed
        return sum(self._a[j] + self._b[j] for j in range(i))

    def alpha(self, i):
        return self._a[i] / self._den

    def beta(self, i):
        return self._b[i] / self._den

    def gamma(self, i):
        return self._g[i] / self._den

    def zeta(self, i):
        return (self._a[i] + self._b[i] - self._g[i]) / self._den

    def eta(self, i):
        return self._eta_numerator(i) / self._den

    def iota(self, i):
        return (self._eta_numerator(i + 1) - self._eta_numerator(i)) / self._eta_numerator(i + 1)

    def iota_alpha(self, i):
        return self._a[i] / self._eta_numerator(i + 1)

    def iota_beta(self, i):
        return self._b[i] / self._eta_numerator(i + 1)

    def fractions(self, which, i):
        """Exact rational value (tests)."""
        num = {"alpha": self._a[i], "beta": self._b[i], "gamma": self._g[i],
               "zeta": self._a[i] + self._b[i] - self._g[i]}[which]
        return Fraction(num, self._den)


# lowstorage.hpp:1242-1306 (Spalart, Moser & Rogers 1991) and :1314-1372 (Yang 2011)
SMR91 = Method("smr91", (29 * (480 // 96), -3 * (480 // 40), 1 * (480 // 6)),
               (37 * (480 // 160), 5 * (480 // 24), 1 * (480 // 6)),
               (8 * (480 // 15), 5 * (480 // 12), 3 * (480 // 4)), 480,
               2.51274532661832862402373, math.sqrt(3.0))
YANG11 = Method("Yang11", (1 * (6 // 3), -1 * (6 // 2), 1 * (6 // 3)), (1 * (6 // 6), 2 * (6 // 3), 0),
                (1 * (6 // 2), 1 * (6 // 3), 1 * (6 // 1)), 6, 2.51274532661832862402373, math.sqrt(3.0))


class State:
    """The state operations of the advance (suzerain/state.hpp: assign_from, add_scaled, exchange) on a
    torch tensor or numpy array held in ``.data``."""

    def __init__(self, data):
        self.data = data

    def assign_from(self, other):
        if hasattr(self.data, "copy_"):
            self.data.copy_(other.data)
        else:
            self.data[...] = other.data

    def add_scaled(self, factor, other):
        if hasattr(self.data, "add_"):
            self.data.add_(other.data, alpha=factor)
        else:
            self.data += factor * other.data

    def exchange(self, other):
        self.data, other.data = other.data, self.data


class InterleavedOperator:
    """Adapter for a device-resident advance: both state buffers keep the interleaved layout
    (npencil, 5, Ny), so that ``exchange`` is a pointer swap (the reference converts between its
    interleaved and contiguous storages there, lowstorage.hpp:1511)."""

    def __init__(self, H, stream=None):
        self.H, self.stream = H, stream

    def apply_mass_plus_scaled_operator(self, phi, state):
        return self.H.apply_mass_plus_scaled_operator(phi, state, stream=self.stream)

    def accumulate_mass_plus_scaled_operator(self, phi, input, beta, output):
        return self.H.accumulate_mass_plus_scaled_operator(phi, input, beta, output, stream=self.stream,
                                                           interleaved_output=True)

    def invert_mass_plus_scaled_operator(self, phi, state):
        return self.H.invert_mass_plus_scaled_operator(phi, state, stream=self.stream)

    def raise_on_singular(self):
        """Fatal singular pencils as in the reference (bsmbsm_solver.cpp:123-141); synchronises."""
        self.H.raise_on_singular()


def delta_t_reducer(candidates):
    """lowstorage::delta_t_reducer: the smallest stable candidate (NaN propagates)."""
    out = math.inf
    for c in candidates:
        if c != c:
            return c
        out = min(out, c)
    return out


def substep(m, L, chi, N, time, a, b, delta_t, substep_index):
    """lowstorage::substep (lowstorage.hpp:1403-1435)."""
    if substep_index >= m.substeps:
        raise ValueError("Requested substep too large")
    L.accumulate_mass_plus_scaled_operator(delta_t * m.alpha(substep_index), a.data,
                                           chi * delta_t * m.zeta(substep_index), b.data)
    N.apply_operator(time + delta_t * m.eta(substep_index), a, m, substep_index)
    b.add_scaled(chi * delta_t * m.gamma(substep_index), a)
    L.invert_mass_plus_scaled_operator(-delta_t * m.beta(substep_index), b.data)
    return delta_t


def step(m, reducer, L, chi, N, time, a, b, max_delta_t=0.0):
    """lowstorage::step (lowstorage.hpp:1471-1520): one apply, two accumulate, three invert for SMR91."""
    b.assign_from(a)
    delta_t = reducer(N.apply_operator(time, b, m, 0))
    if max_delta_t > 0:
        delta_t = delta_t if delta_t != delta_t else min(delta_t, max_delta_t)     # math::minnan
    L.apply_mass_plus_scaled_operator(delta_t * m.alpha(0), a.data)
    a.add_scaled(chi * delta_t * m.gamma(0), b)
    L.invert_mass_plus_scaled_operator(-delta_t * m.beta(0), a.data)
    for i in range(1, m.substeps):
        L.accumulate_mass_plus_scaled_operator(delta_t * m.alpha(i), a.data, chi * delta_t * m.zeta(i), b.data)
        b.exchange(a)
        N.apply_operator(time + delta_t * m.eta(i), b, m, i)
        a.add_scaled(chi * delta_t * m.gamma(i), b)
        L.invert_mass_plus_scaled_operator(-delta_t * m.beta(i), a.data)
    # a device-resident operator records zgbtrf's info per pencil instead of stopping mid-step: make a singular
    # operator fatal here as it is in the reference (bsmbsm_solver.cpp:123-141); one synchronisation per step
    if hasattr(L, "raise_on_singular"):
        L.raise_on_singular()
    return delta_t
