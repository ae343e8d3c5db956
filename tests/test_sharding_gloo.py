"""Multi-rank (kx,kz) sharding on CPU: two gloo ranks each own a block of kz rows, run the
per-pencil path on their block (the oracle stands in for the device kernels here: the
subject of the test is the host-side partition), and the gathered state must equal the
single-rank result -- the implicit operator needs no data-path collective."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import suzerain_b200 as sz
    from suzerain_b200 import shard, synth
    import parity_common as pc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Nx, Ny, Nz, k, htdelta, _ = synth.CONFIGS["tiny_16x24x16"]
        case = pc.make_case("tiny_16x24x16")                       # operators, refs, phi
        g = sz.wavegrid(Nx, Nz, synth.LX, synth.LZ)
        km, kn, act = sz.wavenumbers(g)
        state = synth.state(km, kn, Ny, synth.SEED).reshape(len(km), -1)      # whole wave space, every rank
        mine = shard.shard_wavegrid(g, rank, world)
        mkm, mkn, mact = sz.wavenumbers(mine)
        nx = g.dkex - g.dkbx
        lo, hi = (mine.dkbz - g.dkbz) * nx, (mine.dkez - g.dkbz) * nx      # my slice of the state
        assert np.array_equal(mkm, km[lo:hi]) and np.array_equal(mkn, kn[lo:hi])
        local = state[lo:hi].copy()
        P = pc.oracle_problem(case)
        solved = P.invert("zgbsv", case.phi, mkm[mact], mkn[mact], local[mact])["x"]
        local[mact] = solved
        local[~mact] = 0                                           # dealiased pencils are zero-filled
        # gather (rank order == state order because shards are contiguous kz blocks)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([local.shape[0]]))
        pad = int(max(s.item() for s in sizes))
        buf = torch.zeros((pad, local.shape[1]), dtype=torch.complex128)
        buf[:local.shape[0]] = torch.from_numpy(local)
        out = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        if rank == 0:
            full = np.concatenate([o.numpy()[:int(s.item())] for o, s in zip(out, sizes)])
            want = state.copy()
            want[act] = P.invert("zgbsv", case.phi, km[act], kn[act], state[act])["x"]
            want[~act] = 0
            err = float(np.abs(full - want).max() / np.abs(want).max())
            zz = shard.owner_of_zero_zero(g, world)
            q.put((full.shape == want.shape, err, zz, [int(s.item()) for s in sizes]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    same_shape, err, zz, sizes = q.get(timeout=10)
    assert same_shape and sum(sizes) == 13 * 24         # (dNx/2+1) * dNz stored pencils (Nx=Nz=16, dN=24)
    assert err <= 1e-14                                  # same arithmetic, only partitioned
    assert zz == 0


def test_cost_weighted_shards_cover_the_grid_and_balance_the_weights():
    """shard_wavegrid with measured row costs (bench.py: row_cost_weights): contiguous blocks that cover every kz
    row exactly once, with block weights within one row of the ideal share."""
    sys.path.insert(0, ROOT)
    import suzerain_b200 as sz
    from suzerain_b200 import shard, synth
    g = sz.wavegrid(64, 48, synth.LX, synth.LZ)
    act = shard.active_rows(g).astype(float)
    n = len(act)
    # rows near the largest |kz| cost 1.6 times the others, as on the big channel grid
    wz = np.array([i if i < g.dNz // 2 + 1 else i - g.dNz for i in range(n)])
    w = act * np.where(np.abs(wz) > 0.6 * (g.Nz // 2), 1.6, 1.0)
    for world in (2, 3, 4, 8):
        rows, sums = [], []
        for r in range(world):
            m = shard.shard_wavegrid(g, r, world, w)
            rows += list(range(m.dkbz, m.dkez))
            sums.append(w[m.dkbz - g.dkbz:m.dkez - g.dkbz].sum())
        assert rows == list(range(g.dkbz, g.dkez))
        assert max(sums) - w.sum() / world <= w.max() + 1e-9
        # the unweighted cut puts more cost on the ranks that own the high wavenumbers
        sums0 = []
        for r in range(world):
            m = shard.shard_wavegrid(g, r, world)
            sums0.append(w[m.dkbz - g.dkbz:m.dkez - g.dkbz].sum())
        assert max(sums) <= max(sums0) + 1e-9


def _refs_worker(rank, world, port_no, q):
    """collect_references over ranks: every rank sums its own planes / z-rows into a zeroed 42 x Ny block,
    all-reduce(SUM), then the chi scaling (apps/perfect/perfect.cpp:1275-1277, 1396-1399).  The oracle stands
    in for the device kernel; the subject is the protocol suzerain_b200.api.collect_references follows."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import port
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Ny, Nz, Nx = 12, 6, 8
        gamma, Ma, alpha, beta = 1.4, 1.5, 0.0, 2.0 / 3.0
        rng = np.random.default_rng(2)
        rho = 1 + 0.2 * rng.uniform(-1, 1, (Ny, Nz, Nx)); u = 0.3 * rng.standard_normal((3, Ny, Nz, Nx))
        p = rho * (1 + 0.2 * rng.uniform(-1, 1, (Ny, Nz, Nx))) / gamma
        s = np.stack([p / (gamma - 1) + Ma * Ma * rho * (u ** 2).sum(axis=0) / 2, rho * u[0], rho * u[1], rho * u[2], rho])
        # rank 0: planes 0..6 (all z); rank 1: planes 7..11 -- and, to exercise partial planes, rank 1 also
        # owns the upper z half of plane 6 while rank 0 only has its lower half
        block = np.zeros((Ny, 42))
        if rank == 0:
            block[:6] = port.collect_references(alpha, beta, gamma, Ma, s[:, :6], False).T
            block[6:7] = port.collect_references(alpha, beta, gamma, Ma, s[:, 6:7, :3], False).T
        else:
            block[7:] = port.collect_references(alpha, beta, gamma, Ma, s[:, 7:], True).T      # owns the top plane
            block[6:7] = port.collect_references(alpha, beta, gamma, Ma, s[:, 6:7, 3:], False).T
        t = torch.from_numpy(block)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.mul_(1.0 / (Nz * Nx))
        if rank == 0:
            want = port.collect_references(alpha, beta, gamma, Ma, s, True).T / (Nz * Nx)
            q.put(float(np.abs(t.numpy() - want).max() / np.abs(want).max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_reference_collection_matches_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_refs_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=10) <= 1e-14
