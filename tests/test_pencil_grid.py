"""pencil_grid replacement: wave <-> physical transforms.

CPU: the numpy oracle (oracle/pencil.py) replays the reference's own tests/test_diffwave_p3dfft.cpp
(analytic derivatives through physical -> wave -> diffwave -> physical, with the reference's
diffwave.c from oracle/_ref), the slab protocol is exchanged between two gloo ranks, and the
FFT library exports what include/suzerain_b200_fft.h declares.
GPU: the CUDA path (transpose kernels + cuFFT through the C ABI) against the oracle, the same
analytic test with the CUDA diffwave, and the multi-rank phases against the single-rank result."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pencil as opencil                         # noqa: E402


# ---------------------------------------------------------------------------
# the synthetic periodic fields of tests/test_diffwave_p3dfft.cpp:100-190 (test_tools.hpp periodic_function)
# ---------------------------------------------------------------------------
def periodic(N, maxmode, shift, L, seed):
    """Real trigonometric polynomial with modes 0..maxmode-1 on N points; returns f(x, nderiv)."""
    rng = np.random.default_rng(seed)
    amp = rng.uniform(0.5, 1.5, maxmode)

    def f(x, d=0):
        out = np.zeros_like(x)
        for k in range(maxmode):
            w = 2 * np.pi * k / L
            # d-th derivative of amp cos(w x + shift)
            out = out + amp[k] * w ** d * np.cos(w * x + shift + d * np.pi / 2)
        return out
    return f


def analytic_case(dNx, Ny, dNz, Nx, Nz, Lx, Lz):
    x = np.arange(dNx) * Lx / dNx
    z = np.arange(dNz) * Lz / dNz
    fx = periodic(dNx, (Nx + 1) // 2, np.pi / 3, Lx, 11)
    fz = periodic(dNz, (Nz + 1) // 2, np.pi / 4, Lz, 17)
    scale_y = (np.arange(Ny) + 1.0)[:, None, None]

    def field(dx=0, dz=0):
        return scale_y * fz(z, dz)[None, :, None] * fx(x, dx)[None, None, :]
    return field


GRID = dict(dNx=24, Ny=7, dNz=18, Nx=16, Nz=12, Lx=4 * np.pi, Lz=4 * np.pi / 3)


def test_oracle_round_trip_and_hermitian_layout():
    g = GRID
    rng = np.random.default_rng(1)
    phys = rng.standard_normal((g["Ny"], g["dNz"], g["dNx"]))
    wave = opencil.physical_to_wave(phys)
    assert wave.shape == (g["dNz"], g["dNx"] // 2 + 1, g["Ny"])
    back = opencil.wave_to_physical(wave, g["dNx"])
    assert np.abs(back - phys * g["dNx"] * g["dNz"]).max() <= 1e-11       # unnormalised pair (chi = 1/(dNx dNz))
    # a single mode lands where diffwave expects it: exp(i (2 pi m x / Lx + 2 pi n z / Lz)) + c.c.
    m, n = 3, -2
    x = np.arange(g["dNx"]) / g["dNx"]
    z = np.arange(g["dNz"]) / g["dNz"]
    one = 2 * np.cos(2 * np.pi * (m * x[None, None, :] + n * z[None, :, None])) * np.ones((g["Ny"], 1, 1))
    w = opencil.physical_to_wave(one) / (g["dNx"] * g["dNz"])
    w[n % g["dNz"], m, :] -= 1.0
    assert np.abs(w).max() <= 1e-13


@pytest.mark.parametrize("dxcnt,dzcnt", [(0, 0), (1, 0), (0, 1), (2, 1)])
def test_oracle_replays_reference_diffwave_p3dfft_test(dxcnt, dzcnt):
    """tests/test_diffwave_p3dfft.cpp with the reference's own diffwave.c (oracle/_ref)."""
    pytest.importorskip("scipy")
    try:
        from oracle import ref
        ref.lib()
    except Exception as e:                                   # noqa: BLE001
        pytest.skip(f"oracle/_ref not built: {e}")
    g = GRID
    field = analytic_case(**g)
    wave = opencil.physical_to_wave(field())
    grid = (g["Nx"], g["dNx"], 0, g["dNx"] // 2 + 1, g["Nz"], g["dNz"], 0, g["dNz"])
    scale = g["dNx"] * g["dNz"]
    out = ref.diffwave(dxcnt, dzcnt, complex(1.0 / scale), wave, g["Lx"], g["Lz"], grid)
    got = opencil.wave_to_physical(out, g["dNx"])
    want = field(dxcnt, dzcnt)
    assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max())


def test_fft_library_exports_every_declared_symbol():
    from suzerain_b200 import pencil
    header = open(os.path.join(ROOT, "include", "suzerain_b200_fft.h")).read()
    declared = set(re.findall(r"\b(szb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(pencil.FFT_PROTOTYPES), declared ^ set(pencil.FFT_PROTOTYPES)
    pencil.load()                                             # binds every prototype or raises
    assert pencil.slab_bounds(10, 4) == opencil.slab_bounds(10, 4) == [0, 2, 5, 7, 10]


def _slab_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import pencil as op
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = GRID
        nxw = g["dNx"] // 2 + 1
        rng = np.random.default_rng(5)
        phys = rng.standard_normal((g["Ny"], g["dNz"], g["dNx"]))            # same on every rank
        wave = op.physical_to_wave(phys)
        zs, ys = op.slab_bounds(g["dNz"], world), op.slab_bounds(g["Ny"], world)
        mine = wave[zs[rank]:zs[rank + 1]]
        # wave -> physical: pack, all-to-all, finish
        def exchange(blocks, shapes):
            """all_to_all_single on the flattened blocks, as suzerain_b200/pencil.py does over NCCL
            (float64 view: gloo has no complex all-to-all)."""
            send = torch.from_numpy(np.concatenate([b.reshape(-1) for b in blocks]).view(np.float64).copy())
            ns = [2 * b.size for b in blocks]
            nr = [2 * int(np.prod(sh)) for sh in shapes]
            recv = torch.zeros(sum(nr), dtype=torch.float64)
            dist.all_to_all_single(recv, send, nr, ns)
            flat = recv.numpy().view(np.complex128)
            out, off = [], 0
            for sh in shapes:
                k = int(np.prod(sh))
                out.append(flat[off:off + k].reshape(sh)); off += k
            return out
        send = op.w2p_pack(mine, ys)
        recv = exchange(send, [(ys[rank + 1] - ys[rank], zs[r + 1] - zs[r], nxw) for r in range(world)])
        got = op.w2p_finish(recv, g["dNx"])
        want = op.wave_to_physical(wave, g["dNx"])[ys[rank]:ys[rank + 1]]
        e1 = float(np.abs(got - want).max() / np.abs(want).max())
        # physical -> wave: start, all-to-all, unpack
        send = op.p2w_start(phys[ys[rank]:ys[rank + 1]], zs)
        recv = exchange(send, [(ys[s + 1] - ys[s], zs[rank + 1] - zs[rank], nxw) for s in range(world)])
        got = op.p2w_unpack(recv)
        e2 = float(np.abs(got - mine).max() / np.abs(mine).max())
        q.put((rank, e1, e2))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_slab_protocol_matches_single_rank():
    """The exchange protocol the CUDA pack kernels implement (block order, counts), on CPU over gloo."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    for _, e1, e2 in res:
        assert e1 <= 1e-13 and e2 <= 1e-13


# ---------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dev():
    import torch
    return torch.device("cuda:0")


def _field_buffer(pg, dev):
    import torch
    return torch.zeros(pg.local_physical_storage(), dtype=torch.float64, device=dev)


@pytest.mark.gpu
@pytest.mark.parametrize("dNx,Ny,dNz", [(24, 7, 18), (288, 96, 288), (30, 33, 45), (2, 1, 1)])
def test_gpu_transforms_match_oracle(dev, dNx, Ny, dNz):
    import torch
    import suzerain_b200 as sz
    pg = sz.PencilGrid(dNx, Ny, dNz, rank=0, nranks=1)
    assert pg.local_wave_extent == (dNx // 2 + 1, Ny, dNz) and pg.local_physical_extent == (dNx, Ny, dNz)
    assert pg.has_zero_zero_modes() and pg.chi() == 1.0 / (dNx * dNz)
    rng = np.random.default_rng(2)
    phys = rng.standard_normal((Ny, dNz, dNx))
    buf = _field_buffer(pg, dev)
    pg.physical_view(buf).copy_(torch.from_numpy(phys))
    pg.transform_physical_to_wave(buf)
    torch.cuda.synchronize()
    want = opencil.physical_to_wave(phys)
    got = pg.wave_view(buf).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    pg.transform_wave_to_physical(buf)
    torch.cuda.synchronize()
    back = pg.physical_view(buf).cpu().numpy()
    assert np.abs(back - phys * dNx * dNz).max() <= 1e-12 * dNx * dNz * np.abs(phys).max()
    # wave -> physical of data that is not a forward transform (imaginary parts in the self-conjugate modes
    # are ignored by complex-to-real, as FFTW / P3DFFT do)
    wave = rng.standard_normal(want.shape) + 1j * rng.standard_normal(want.shape)
    pg.wave_view(buf).copy_(torch.from_numpy(wave))
    pg.transform_wave_to_physical(buf)
    torch.cuda.synchronize()
    w0 = opencil.wave_to_physical(wave, dNx)
    assert np.abs(pg.physical_view(buf).cpu().numpy() - w0).max() <= 1e-12 * np.abs(w0).max()


@pytest.mark.gpu
@pytest.mark.parametrize("dxcnt,dzcnt", [(1, 0), (0, 1), (2, 1)])
def test_gpu_replays_reference_diffwave_p3dfft_test(dev, dxcnt, dzcnt):
    """tests/test_diffwave_p3dfft.cpp on the device: transform, CUDA diffwave, transform back."""
    import torch
    import suzerain_b200 as sz
    g = GRID
    field = analytic_case(**g)
    pg = sz.PencilGrid(g["dNx"], g["Ny"], g["dNz"], rank=0, nranks=1)
    buf = _field_buffer(pg, dev)
    pg.physical_view(buf).copy_(torch.from_numpy(field()))
    pg.transform_physical_to_wave(buf)
    wg = sz.wavegrid(g["Nx"], g["Nz"], g["Lx"], g["Lz"])
    assert (wg.dNx, wg.dNz) == (g["dNx"], g["dNz"])
    sz.diffwave_apply(dxcnt, dzcnt, 1.0 / (g["dNx"] * g["dNz"]), pg.wave_view(buf), wg)
    pg.transform_wave_to_physical(buf)
    torch.cuda.synchronize()
    want = field(dxcnt, dzcnt)
    got = pg.physical_view(buf).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max())


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_gpu_slab_phases_compose_to_the_single_rank_transform(dev, world):
    """The multi-rank phases (pack / finish / start / unpack) of every rank, run on one GPU with the
    all-to-all done by hand, against the oracle's blocks and the whole-field transform."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import pencil
    import ctypes as C
    g = dict(dNx=30, Ny=11, dNz=20)
    nxw = g["dNx"] // 2 + 1
    rng = np.random.default_rng(3)
    phys = rng.standard_normal((g["Ny"], g["dNz"], g["dNx"]))
    wave = opencil.physical_to_wave(phys)
    zs, ys = opencil.slab_bounds(g["dNz"], world), opencil.slab_bounds(g["Ny"], world)
    L = pencil.load()
    grids = [sz.PencilGrid(g["dNx"], g["Ny"], g["dNz"], rank=r, nranks=world) for r in range(world)]
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    # ---- wave -> physical ----
    sends = []
    for r, pg in enumerate(grids):
        assert pg.local_wave_start[2] == zs[r] and pg.local_physical_start[1] == ys[r]
        ns, nr = pg._counts(0)
        assert ns == [(ys[q + 1] - ys[q]) * (zs[r + 1] - zs[r]) * nxw for q in range(world)]
        w = torch.from_numpy(np.ascontiguousarray(wave[zs[r]:zs[r + 1]])).to(dev)
        send = torch.zeros(sum(ns), dtype=torch.complex128, device=dev)
        assert L.szb_pencil_grid_w2p_pack(C.c_void_p(pg._h), ptr(w), ptr(send), s) == 0
        blocks = list(torch.split(send, ns))
        for q, b in enumerate(opencil.w2p_pack(wave[zs[r]:zs[r + 1]], ys)):
            assert np.array_equal(blocks[q].cpu().numpy().reshape(b.shape), b)
        sends.append(blocks)
    want = opencil.wave_to_physical(wave, g["dNx"])
    for r, pg in enumerate(grids):
        recv = torch.cat([sends[q][r] for q in range(world)])
        out = torch.zeros((ys[r + 1] - ys[r], g["dNz"], g["dNx"]), dtype=torch.float64, device=dev)
        assert L.szb_pencil_grid_w2p_finish(C.c_void_p(pg._h), ptr(recv), ptr(out), s) == 0
        torch.cuda.synchronize()
        assert np.abs(out.cpu().numpy() - want[ys[r]:ys[r + 1]]).max() <= 1e-12 * np.abs(want).max()
    # ---- physical -> wave ----
    sends = []
    for r, pg in enumerate(grids):
        ns, nr = pg._counts(1)
        p = torch.from_numpy(np.ascontiguousarray(phys[ys[r]:ys[r + 1]])).to(dev)
        send = torch.zeros(sum(ns), dtype=torch.complex128, device=dev)
        assert L.szb_pencil_grid_p2w_start(C.c_void_p(pg._h), ptr(p), ptr(send), s) == 0
        sends.append(list(torch.split(send, ns)))
    for r, pg in enumerate(grids):
        recv = torch.cat([sends[q][r] for q in range(world)])
        out = torch.zeros((zs[r + 1] - zs[r], nxw, g["Ny"]), dtype=torch.complex128, device=dev)
        assert L.szb_pencil_grid_p2w_unpack(C.c_void_p(pg._h), ptr(recv), ptr(out), s) == 0
        torch.cuda.synchronize()
        assert np.abs(out.cpu().numpy() - wave[zs[r]:zs[r + 1]]).max() <= 1e-12 * np.abs(wave).max()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_gpu_peer_memory_phases_compose_to_the_single_rank_transform(dev, world):
    """The peer-memory variants (the transposing kernels store straight into every rank's FFT / wave
    buffer).  All "ranks" live on one GPU here, so the peer addresses are ordinary device pointers."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import pencil
    import ctypes as C
    g = dict(dNx=30, Ny=11, dNz=20)
    nxw = g["dNx"] // 2 + 1
    rng = np.random.default_rng(4)
    phys = rng.standard_normal((g["Ny"], g["dNz"], g["dNx"]))
    wave = opencil.physical_to_wave(phys)
    zs, ys = opencil.slab_bounds(g["dNz"], world), opencil.slab_bounds(g["Ny"], world)
    L = pencil.load()
    grids = [sz.PencilGrid(g["dNx"], g["Ny"], g["dNz"], rank=r, nranks=world) for r in range(world)]
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    fft = [torch.full((ys[r + 1] - ys[r], g["dNz"], nxw), float("nan"), dtype=torch.complex128, device=dev) for r in range(world)]
    wav = [torch.full((zs[r + 1] - zs[r], nxw, g["Ny"]), float("nan"), dtype=torch.complex128, device=dev) for r in range(world)]
    pf = (C.c_ulonglong * world)(*[t.data_ptr() for t in fft])
    pw = (C.c_ulonglong * world)(*[t.data_ptr() for t in wav])
    # wave -> physical
    for r, pg in enumerate(grids):
        w = torch.from_numpy(np.ascontiguousarray(wave[zs[r]:zs[r + 1]])).to(dev)
        assert L.szb_pencil_grid_w2p_pack_peers(C.c_void_p(pg._h), ptr(w), pf, s) == 0
    torch.cuda.synchronize()
    want = opencil.wave_to_physical(wave, g["dNx"])
    for r, pg in enumerate(grids):
        assert np.array_equal(fft[r].cpu().numpy(), np.transpose(wave, (2, 0, 1))[ys[r]:ys[r + 1]])
        out = torch.zeros((ys[r + 1] - ys[r], g["dNz"], g["dNx"]), dtype=torch.float64, device=dev)
        assert L.szb_pencil_grid_w2p_fft(C.c_void_p(pg._h), ptr(fft[r]), ptr(out), s) == 0
        torch.cuda.synchronize()
        assert np.abs(out.cpu().numpy() - want[ys[r]:ys[r + 1]]).max() <= 1e-12 * np.abs(want).max()
    # physical -> wave
    for r, pg in enumerate(grids):
        p = torch.from_numpy(np.ascontiguousarray(phys[ys[r]:ys[r + 1]])).to(dev)
        assert L.szb_pencil_grid_p2w_fft(C.c_void_p(pg._h), ptr(p), ptr(fft[r]), s) == 0
        assert L.szb_pencil_grid_p2w_scatter_peers(C.c_void_p(pg._h), ptr(fft[r]), pw, s) == 0
    torch.cuda.synchronize()
    for r in range(world):
        assert np.abs(wav[r].cpu().numpy() - wave[zs[r]:zs[r + 1]]).max() <= 1e-12 * np.abs(wave).max()
