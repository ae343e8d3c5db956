"""(kx,kz) sharding of wave space across ranks.

The reference decomposes wave space with a 2-D pencil grid in which the wall-normal
direction stays complete on every rank (suzerain/pencil_grid.cpp:118-157), so each
rank owns whole pencils and the implicit operator never communicates
(SURVEY.md 8e).  Here every rank gets a contiguous block of kz rows (z is the
slowest state index, suzerain/storage.hpp:235-236), chosen so that the number of
ACTIVE (non-dealiased) pencils -- the ones that cost a banded solve -- is balanced;
dealiased pencils only cost a memset.
"""
from __future__ import annotations

import numpy as np

from . import lib as _L


def _wavenumber(N, i):
    return i if i < N // 2 + 1 else -N + i          # suzerain/inorder.h:92-96


def active_rows(grid):
    """Active-pencil count of every local kz row of `grid`."""
    kx_active = sum(1 for m in range(grid.dkbx, grid.dkex)
                    if abs(_wavenumber(grid.dNx, m)) <= (grid.Nx - 1) // 2)
    return np.array([kx_active if abs(_wavenumber(grid.dNz, n)) <= (grid.Nz - 1) // 2 else 0
                     for n in range(grid.dkbz, grid.dkez)], dtype=np.int64)


def shard_bounds(weights, world):
    """Splits len(weights) rows into `world` contiguous blocks with balanced weight sums;
    returns world+1 boundaries.  Every block is non-empty when there are enough rows."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cum, target, side="left"))
        # choose the closer of b-1 / b to the target, keep blocks non-empty and ordered
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[min(b, n)] - target):
            b -= 1
        b = max(b, bounds[-1] + (1 if n >= world else 0))
        b = min(b, n - (world - r) if n >= world else n)
        bounds.append(b)
    bounds.append(n)
    return bounds


def shard_wavegrid(grid, rank, world, weights=None):
    """The sub-grid (same global extents, narrower local kz range) owned by `rank`.  `weights`: cost of every
    local kz row (default: its number of active pencils).  The banded solve of a pencil costs more where its
    pivot sequence leaves the diagonal -- up to 29 % of the panels at the highest wavenumbers of the
    1536x384x1152 grid against 3 % elsewhere -- so a stepper may pass measured row costs instead
    (bench.py: row_cost_weights); every rank must use the same weights."""
    b = shard_bounds(active_rows(grid) if weights is None else weights, world)
    return _L.WaveGrid(grid.Nx, grid.dNx, grid.dkbx, grid.dkex, grid.Nz, grid.dNz,
                       grid.dkbz + b[rank], grid.dkbz + b[rank + 1], grid.Lx, grid.Lz)


def owner_of_zero_zero(grid, world):
    """Rank owning the (0,0) pencil (it also solves the integral-constraint columns,
    apps/perfect/operator_hybrid_isothermal.cpp:676-685); -1 if not in `grid`."""
    if not (grid.dkbx <= 0 < grid.dkex and grid.dkbz <= 0 < grid.dkez):
        return -1
    b = shard_bounds(active_rows(grid), world)
    row = 0 - grid.dkbz
    for r in range(world):
        if b[r] <= row < b[r + 1]:
            return r
    return -1
