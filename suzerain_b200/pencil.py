"""Device replacement of ``suzerain::pencil_grid`` (suzerain/pencil_grid.hpp:57-246): the wave <->
physical transforms of the nonlinear operator, behind ``libsuzerain_b200_fft.so``
(include/suzerain_b200_fft.h).  One process per GPU; wave space is cut in Z, physical space in Y,
and with more than one rank a transform is  pack -> NCCL all-to-all -> finish.

Layouts (the reference's): wave = complex ``[Z][X][Y]`` with Y fastest and X = dNx/2+1,
physical = real ``[Y][Z][X]`` with X fastest; transforms are unnormalised (a round trip
multiplies by dNx*dNz = 1/chi)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
FFT_LIB_PATH = os.path.join(_HERE, "libsuzerain_b200_fft.so")
_lib = None

c_void_p, I3, LLP = C.c_void_p, C.c_int * 3, C.POINTER(C.c_longlong)
FFT_PROTOTYPES = {
    "szb_pencil_grid_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(c_void_p)]),
    "szb_pencil_grid_destroy": (None, [c_void_p]),
    "szb_pencil_grid_extents": (C.c_int, [c_void_p, C.c_int, C.c_int, I3, I3]),
    "szb_pencil_grid_local_wave_storage": (C.c_size_t, [c_void_p]),
    "szb_pencil_grid_local_physical_storage": (C.c_size_t, [c_void_p]),
    "szb_pencil_grid_has_zero_zero_modes": (C.c_int, [c_void_p]),
    "szb_pencil_grid_transform_wave_to_physical": (C.c_int, [c_void_p, c_void_p, c_void_p]),
    "szb_pencil_grid_transform_physical_to_wave": (C.c_int, [c_void_p, c_void_p, c_void_p]),
    "szb_pencil_grid_exchange_counts": (C.c_int, [c_void_p, C.c_int, LLP, LLP]),
    "szb_pencil_grid_w2p_pack": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "szb_pencil_grid_w2p_finish": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "szb_pencil_grid_p2w_start": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "szb_pencil_grid_p2w_unpack": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "szb_pencil_grid_w2p_pack_peers": (C.c_int, [c_void_p, c_void_p, C.POINTER(C.c_ulonglong), c_void_p]),
    "szb_pencil_grid_w2p_fft": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "szb_pencil_grid_p2w_fft": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "szb_pencil_grid_p2w_scatter_peers": (C.c_int, [c_void_p, c_void_p, C.POINTER(C.c_ulonglong), c_void_p]),
    "szb_fft_launch_count": (C.c_ulonglong, []),
}


def load():
    """Load libsuzerain_b200_fft.so; raises OSError if it has not been built (no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(FFT_LIB_PATH):
            raise OSError(f"{FFT_LIB_PATH} is missing: build it with `make -C suzerain_b200/csrc`.  "
                          "There is no CPU fallback.")
        lib = C.CDLL(FFT_LIB_PATH)
        for name, (res, args) in FFT_PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def slab_bounds(n, nranks):
    """Contiguous balanced cut of 0..n over the ranks: rank r owns [r n / R, (r+1) n / R)."""
    return [r * n // nranks for r in range(nranks + 1)]


class PencilGrid:
    """``suzerain::pencil_grid`` on the device.  ``group`` is a torch.distributed process group (or
    None for a single process); the tensors passed to the transforms live on this rank's GPU."""

    def __init__(self, dNx, Ny, dNz, group=None, rank=None, nranks=None, exchange="auto"):
        """exchange: "p2p" (the transposing kernels store straight into the peers' buffers over NVLink;
        needs torch symmetric memory), "nccl" (pack -> all_to_all_single -> finish), or "auto": p2p when
        the symmetric buffers can be set up on every rank, else nccl."""
        import torch.distributed as dist
        self.group = group
        if nranks is None:
            nranks = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
            rank = dist.get_rank(group) if nranks > 1 else 0
        self.rank, self.nranks = int(rank), int(nranks)
        self.global_physical_extent = (dNx, Ny, dNz)
        self.global_wave_extent = (dNx // 2 + 1, Ny, dNz)
        h = c_void_p()
        rc = load().szb_pencil_grid_create(dNx, Ny, dNz, self.nranks, self.rank, C.byref(h))
        if rc:
            raise RuntimeError(f"szb_pencil_grid_create failed: {rc}")
        self._h = h.value
        self.local_physical_start, self.local_physical_end = self._extents(0, self.rank)
        self.local_wave_start, self.local_wave_end = self._extents(1, self.rank)
        self.local_physical_extent = tuple(e - s for s, e in zip(self.local_physical_start, self.local_physical_end))
        self.local_wave_extent = tuple(e - s for s, e in zip(self.local_wave_start, self.local_wave_end))
        self._bufs = {}
        assert exchange in ("auto", "nccl", "p2p")
        self.exchange = exchange if self.nranks > 1 else "local"
        self._p2p = None

    def _resolve_exchange(self, device):
        """"auto": every rank tries to set up the symmetric buffers; all must succeed."""
        if self.exchange != "auto":
            return
        import torch
        import torch.distributed as dist
        ok = 1
        try:
            self._p2p_setup(device)
        except Exception:                                    # noqa: BLE001
            ok, self._p2p = 0, None
        t = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        self.exchange = "p2p" if int(t.item()) == 1 else "nccl"

    def _p2p_setup(self, device, nfields=1):
        """Symmetric buffers of every rank, per field: fft = [max Yloc][dNz][X], wave = [max Zloc][X][Y]
        complex.  Re-made (collectively) when more fields per exchange are asked for than last time."""
        if self._p2p is not None and self._p2p["nf"] < nfields:
            self._p2p = None
        if self._p2p is None:
            import torch
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            nxw, Ny, dNz = self.global_wave_extent
            R = self.nranks
            ymax = max((r + 1) * Ny // R - r * Ny // R for r in range(R))
            zmax = max((r + 1) * dNz // R - r * dNz // R for r in range(R))
            group = self.group if self.group is not None else dist.group.WORLD
            nf = max(1, int(nfields))
            sf, sw = 2 * ymax * dNz * nxw, 2 * zmax * nxw * Ny           # doubles per field
            fft = symm.empty(nf * sf, dtype=torch.float64, device=device)
            wave = symm.empty(nf * sw, dtype=torch.float64, device=device)
            hf, hw = symm.rendezvous(fft, group), symm.rendezvous(wave, group)
            pf = [(C.c_ulonglong * R)(*[int(p) + 8 * f * sf for p in hf.buffer_ptrs]) for f in range(nf)]
            pw = [(C.c_ulonglong * R)(*[int(p) + 8 * f * sw for p in hw.buffer_ptrs]) for f in range(nf)]
            self._p2p = dict(nf=nf, sf=sf, sw=sw, fft=fft, wave=wave, hf=hf, hw=hw, pf=pf, pw=pw)
        return self._p2p

    def __del__(self):
        try:
            if self._h:
                load().szb_pencil_grid_destroy(c_void_p(self._h))
                self._h = None
        except Exception:
            pass

    def _extents(self, which, r):
        s, e = I3(), I3()
        load().szb_pencil_grid_extents(c_void_p(self._h), which, r, s, e)
        return tuple(s), tuple(e)

    def chi(self):
        """pencil_grid::chi (pencil_grid.hpp:159-163)."""
        return 1.0 / (self.global_physical_extent[0] * self.global_physical_extent[2])

    def has_zero_zero_modes(self):
        return bool(load().szb_pencil_grid_has_zero_zero_modes(c_void_p(self._h)))

    def local_wave_storage(self):
        return load().szb_pencil_grid_local_wave_storage(c_void_p(self._h))

    def local_physical_storage(self):
        return load().szb_pencil_grid_local_physical_storage(c_void_p(self._h))

    # ---- views of one field's storage (a float64 tensor of local_physical_storage() elements) ----
    def wave_view(self, buf):
        import torch
        nx, ny, nz = self.local_wave_extent
        return torch.view_as_complex(buf[:2 * nx * ny * nz].view(-1, 2)).view(nz, nx, ny)

    def physical_view(self, buf):
        nx, ny, nz = self.local_physical_extent
        return buf[:nx * ny * nz].view(ny, nz, nx)

    def _counts(self, direction):
        send = (C.c_longlong * self.nranks)()
        recv = (C.c_longlong * self.nranks)()
        load().szb_pencil_grid_exchange_counts(c_void_p(self._h), direction, send, recv)
        return list(send), list(recv)

    def _exchange_buffers(self, direction, device):
        import torch
        key = (direction, str(device))
        if key not in self._bufs:
            send, recv = self._counts(direction)
            self._bufs[key] = (torch.empty(max(1, sum(send)), dtype=torch.complex128, device=device),
                               torch.empty(max(1, sum(recv)), dtype=torch.complex128, device=device), send, recv)
        return self._bufs[key]

    def _all_to_all(self, recv, send, nr, ns):
        """One block per peer, in rank order (NCCL; the complex buffers travel as pairs of doubles)."""
        import torch
        import torch.distributed as dist
        dist.all_to_all_single(torch.view_as_real(recv[:sum(nr)]).reshape(-1), torch.view_as_real(send[:sum(ns)]).reshape(-1),
                               [2 * c for c in nr], [2 * c for c in ns], group=self.group)

    def _stream(self, stream):
        import torch
        return c_void_p((stream or torch.cuda.current_stream()).cuda_stream)

    def _check(self, buf):
        import torch
        assert buf.is_cuda and buf.dtype == torch.float64 and buf.is_contiguous()
        assert buf.numel() >= self.local_physical_storage()

    def transform_wave_to_physical(self, buf, stream=None):
        """pencil_grid::transform_wave_to_physical (pencil_grid.hpp:200-205), in place."""
        self.transform_wave_to_physical_many([buf], stream)

    def transform_physical_to_wave(self, buf, stream=None):
        """pencil_grid::transform_physical_to_wave (pencil_grid.hpp:216-221), in place."""
        self.transform_physical_to_wave_many([buf], stream)

    def transform_wave_to_physical_many(self, bufs, stream=None):
        """Several fields at once (the nonlinear operator transforms 5 + 12 of them per substep,
        navier_stokes.hpp:320-331): with the peer-memory exchange the fields share one pair of barriers."""
        for b in bufs:
            self._check(b)
        dev = bufs[0].device
        self._resolve_exchange(dev)
        import torch
        ts = stream or torch.cuda.current_stream(dev)
        L, h, s = load(), c_void_p(self._h), c_void_p(ts.cuda_stream)
        rc = 0
        if self.nranks == 1:
            for b in bufs:
                rc = rc or L.szb_pencil_grid_transform_wave_to_physical(h, c_void_p(b.data_ptr()), s)
        else:
            # the symmetric-memory barriers, NCCL and torch copies are issued on torch's CURRENT stream: make the
            # caller's stream current so that kernels, barriers and the exchange are ordered on one stream.  Every
            # barrier / collective is executed even after a failed launch: the peers are waiting in theirs.
            with torch.cuda.stream(ts):
                if self.exchange == "p2p":
                    P = self._p2p_setup(dev, len(bufs))
                    P["hf"].barrier(channel=0)                      # every peer's FFT buffers are free again
                    for f, b in enumerate(bufs):
                        rc = rc or L.szb_pencil_grid_w2p_pack_peers(h, c_void_p(b.data_ptr()), P["pf"][f], s)
                    P["hf"].barrier(channel=1)                      # every block has landed
                    for f, b in enumerate(bufs):
                        rc = rc or L.szb_pencil_grid_w2p_fft(h, c_void_p(P["fft"].data_ptr() + 8 * f * P["sf"]),
                                                             c_void_p(b.data_ptr()), s)
                else:
                    send, recv, ns, nr = self._exchange_buffers(0, dev)
                    for b in bufs:
                        rc = rc or L.szb_pencil_grid_w2p_pack(h, c_void_p(b.data_ptr()), c_void_p(send.data_ptr()), s)
                        self._all_to_all(recv, send, nr, ns)
                        rc = rc or L.szb_pencil_grid_w2p_finish(h, c_void_p(recv.data_ptr()), c_void_p(b.data_ptr()), s)
        if rc:
            raise RuntimeError(f"transform_wave_to_physical failed: {rc}")

    def transform_physical_to_wave_many(self, bufs, stream=None):
        for b in bufs:
            self._check(b)
        dev = bufs[0].device
        self._resolve_exchange(dev)
        import torch
        ts = stream or torch.cuda.current_stream(dev)
        L, h, s = load(), c_void_p(self._h), c_void_p(ts.cuda_stream)
        rc = 0
        if self.nranks == 1:
            for b in bufs:
                rc = rc or L.szb_pencil_grid_transform_physical_to_wave(h, c_void_p(b.data_ptr()), s)
        else:
            with torch.cuda.stream(ts):                             # see transform_wave_to_physical_many
                if self.exchange == "p2p":
                    P = self._p2p_setup(dev, len(bufs))
                    for f, b in enumerate(bufs):
                        rc = rc or L.szb_pencil_grid_p2w_fft(h, c_void_p(b.data_ptr()), c_void_p(P["fft"].data_ptr() + 8 * f * P["sf"]), s)
                    P["hw"].barrier(channel=0)                      # every peer's wave buffers are free again
                    for f, b in enumerate(bufs):
                        rc = rc or L.szb_pencil_grid_p2w_scatter_peers(h, c_void_p(P["fft"].data_ptr() + 8 * f * P["sf"]), P["pw"][f], s)
                    P["hw"].barrier(channel=1)
                    nx, ny, nz = self.local_wave_extent
                    for f, b in enumerate(bufs):
                        b[:2 * nx * ny * nz].copy_(P["wave"][f * P["sw"]:f * P["sw"] + 2 * nx * ny * nz])
                else:
                    send, recv, ns, nr = self._exchange_buffers(1, dev)
                    for b in bufs:
                        rc = rc or L.szb_pencil_grid_p2w_start(h, c_void_p(b.data_ptr()), c_void_p(send.data_ptr()), s)
                        self._all_to_all(recv, send, nr, ns)
                        rc = rc or L.szb_pencil_grid_p2w_unpack(h, c_void_p(recv.data_ptr()), c_void_p(b.data_ptr()), s)
        if rc:
            raise RuntimeError(f"transform_physical_to_wave failed: {rc}")
