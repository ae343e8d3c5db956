"""Host-side mirror of the reference's operator interface over the C ABI.

Names follow the reference: ``BsplineOp`` ~ ``suzerain::bsplineop``
(suzerain/bspline.hpp:348-590), ``ImexOp`` ~ the per-pencil
``suzerain_rholut_imexop_*`` family (suzerain/rholut_imexop.h) bound to a
scenario / reference profiles / wall data, and ``OperatorHybridIsothermal`` ~
``suzerain::perfect::operator_hybrid_isothermal``
(apps/perfect/operator_hybrid_isothermal.hpp:101-140) with its three
``*_mass_plus_scaled_operator`` methods.

Device arrays are ``torch`` CUDA tensors (torch is only the allocator / stream
provider here); every computation is a call into libsuzerain_b200.so.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import lib as _L


def _d2(z):
    z = complex(z)
    return (C.c_double * 2)(z.real, z.imag)


def _ptr(t):
    """Raw device/host pointer of a torch tensor or numpy array (or None)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def _stream_handle(stream):
    if stream is None:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


class BsplineOp:
    """B-spline Greville-collocation operators D^(0..nderiv) (host)."""

    def __init__(self, handle):
        self._h = handle
        L = _L.load()
        h = C.c_void_p(handle)
        self.k = L.szb_bsplineop_k(h)
        self.n = L.szb_bsplineop_n(h)
        self.nderiv = L.szb_bsplineop_nderiv(h)
        self.kl = np.array([L.szb_bsplineop_kl(h, d) for d in range(self.nderiv + 1)], dtype=np.int32)
        self.ku = np.array([L.szb_bsplineop_ku(h, d) for d in range(self.nderiv + 1)], dtype=np.int32)
        self.max_kl = L.szb_bsplineop_max_kl(h)
        self.max_ku = L.szb_bsplineop_max_ku(h)
        self.ld = L.szb_bsplineop_ld(h)
        self.breakpoints = None

    @classmethod
    def from_breakpoints(cls, k, breakpoints, nderiv=None):
        L = _L.load()
        b = np.ascontiguousarray(breakpoints, dtype=np.float64)
        if nderiv is None:
            nderiv = max(k - 2, 0)           # suzerain/support/support.cpp:308
        h = C.c_void_p()
        _L.check("szb_bsplineop_alloc",
                 L.szb_bsplineop_alloc(k, len(b), b.ctypes.data_as(_L.c_double_p), nderiv, C.byref(h)))
        out = cls(h.value)
        out.breakpoints = b.copy()
        return out

    def knots(self):
        """The clamped knot vector: each end breakpoint with multiplicity k (what the reference stores as
        /knots next to /breakpoints_y, support.cpp save_bsplines)."""
        b = self.breakpoints
        return np.concatenate([np.full(self.k - 1, b[0]), b, np.full(self.k - 1, b[-1])])

    def integration_weights(self):
        """w with  int f dy = w . coefficients  (suzerain_bspline_integration_coefficients, used by the bulk
        constraints of treatment_constraint.cpp):  int B_j = (t_{j+k} - t_j) / k."""
        t = self.knots()
        return (t[self.k:] - t[:-self.k]) / self.k

    @classmethod
    def from_storage(cls, k, n, nderiv, kl, ku, storage):
        L = _L.load()
        kl = np.ascontiguousarray(kl, dtype=np.int32)
        ku = np.ascontiguousarray(ku, dtype=np.int32)
        st = np.ascontiguousarray(storage, dtype=np.float64)
        h = C.c_void_p()
        _L.check("szb_bsplineop_from_storage",
                 L.szb_bsplineop_from_storage(k, n, nderiv, kl.ctypes.data_as(_L.c_int_p),
                                              ku.ctypes.data_as(_L.c_int_p),
                                              st.ctypes.data_as(_L.c_double_p), C.byref(h)))
        return cls(h.value)

    def __del__(self):
        try:
            if self._h:
                _L.load().szb_bsplineop_free(C.c_void_p(self._h))
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return C.c_void_p(self._h)

    @property
    def storage(self):
        """(nderiv+1, n, ld) copy of the operator block in the reference's layout:
        storage[d, i, max_ku + j - i] = D^(d)[i, j]."""
        L = _L.load()
        out = np.empty((self.nderiv + 1, self.n, self.ld))
        for d in range(self.nderiv + 1):
            p = L.szb_bsplineop_D_T(self.handle, d)
            addr = C.addressof(p.contents) - 8 * int(self.max_ku - self.ku[d])
            buf = (C.c_double * (self.n * self.ld)).from_address(addr)
            out[d] = np.frombuffer(buf, dtype=np.float64).reshape(self.n, self.ld)
        return out

    def D_T_offset(self, d):
        return int(self.max_ku - self.ku[d])

    def greville(self):
        xi = np.empty(self.n)
        _L.check("szb_bsplineop_greville",
                 _L.load().szb_bsplineop_greville(self.handle, xi.ctypes.data_as(_L.c_double_p)))
        return xi

    def dense(self, d):
        st = self.storage[d]
        out = np.zeros((self.n, self.n))
        for i in range(self.n):
            for j in range(max(0, i - self.max_ku), min(self.n, i + self.max_kl + 1)):
                out[i, j] = st[i, self.max_ku + j - i]
        return out


def htstretch_breakpoints(ndof, k, left, right, htdelta):
    """suzerain/support/support.cpp:288-300."""
    L = _L.load()
    x = np.linspace(0.0, 1.0, ndof - k + 2)
    if htdelta >= 0:
        b = np.array([L.szb_htstretch2(+htdelta, 1.0, v) for v in x])
    else:
        b = np.array([L.szb_htstretch1(-htdelta, 1.0, v) for v in x])
    return (right - left) * b + left


def _parse_bool(text):
    t = text.strip().lower()
    if t in ("true", "1"):
        return True
    if t in ("false", "0"):
        return False
    raise ValueError(f"not a boolean: {text!r}")


@dataclasses.dataclass
class SolverSpec:
    """specification_zgbsv (suzerain/specification_zgbsv.cpp:46-124).

    ``reuse`` and ``siter`` are accepted for drop-in compatibility and are hints on the
    device: every pencil is factored afresh in double precision and refined to the same
    stopping criterion (the reference's own test holds the variants to 6e-15 of each other,
    apps/perfect/test_implicit_solvers.sh:24-50)."""
    method: str = "zcgbsvx"
    aiter: int = 1
    diter: int = 5
    tolsc: float = 0.0
    equil: bool = False
    reuse: bool = False
    siter: int = -1

    @classmethod
    def parse(cls, text: str):
        """The reference grammar: zgbsv | zgbsvx[,equil=b] | zcgbsvx[,reuse=b][,aiter=i][,siter=i]
        [,diter=i][,tolsc=d]; case-insensitive, whitespace ignored, empty = defaults."""
        parts = [p.strip() for p in text.split(",")]
        if parts == [""]:
            return cls()
        head = parts[0].lower()
        if head not in ("zgbsv", "zgbsvx", "zcgbsvx") or any(p == "" for p in parts):
            raise ValueError(f"spec_zgbsv specification {text!r} invalid")
        spec = cls(method=head)
        allowed = {"zgbsv": (), "zgbsvx": ("equil",), "zcgbsvx": ("reuse", "aiter", "siter", "diter", "tolsc")}[head]
        seen = set()
        for kv in parts[1:]:
            key, eq, val = kv.partition("=")
            key = key.strip().lower()
            if not eq or key not in allowed or key in seen:
                raise ValueError(f"spec_zgbsv specification {text!r} invalid beginning with {kv!r}")
            seen.add(key)
            if key in ("equil", "reuse"):
                setattr(spec, key, _parse_bool(val))
            elif key == "tolsc":
                spec.tolsc = float(val)
            else:
                setattr(spec, key, int(val))
        return spec

    def in_place(self):
        """specification_zgbsv::in_place (specification_zgbsv.cpp:116-124)."""
        return self.method == "zgbsv"

    def c(self):
        return _L.ZgbsvSpec({"zgbsv": 0, "zcgbsvx": 1, "zgbsvx": 2}[self.method], self.aiter, self.diter,
                            self.tolsc, int(self.equil), int(self.reuse), self.siter)


class ImexOp:
    """Device-resident (M + phi L) for one scenario / set of reference profiles."""

    def __init__(self, bop: BsplineOp):
        self.bop = bop
        h = C.c_void_p()
        _L.check("szb_imexop_create", _L.load().szb_imexop_create(bop.handle, C.byref(h)))
        self._h = h.value
        A = _L.load().szb_imexop_bsmbsm(self.handle)
        self.S, self.n, self.N, self.KL, self.KU, self.LD = A.S, A.n, A.N, A.KL, A.KU, A.LD
        self._keep = []

    def __del__(self):
        try:
            if self._h:
                _L.load().szb_imexop_destroy(C.c_void_p(self._h))
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return C.c_void_p(self._h)

    def set_scenario(self, Re, Pr, Ma, alpha, gamma):
        s = _L.Scenario(Re, Pr, Ma, alpha, gamma)
        _L.check("szb_imexop_set_scenario", _L.load().szb_imexop_set_scenario(self.handle, C.byref(s)))
        return self

    def set_refs(self, refs, stride=None):
        """refs: (26, n) array in lib.REF_NAMES order, or a (n, 42)-style array
        plus explicit row map via ``stride`` is not needed here: rows are dense."""
        refs = np.ascontiguousarray(refs, dtype=np.float64)
        assert refs.shape == (26, self.n)
        r, ld = _L.Ref(), _L.RefLd()
        for q, name in enumerate(_L.REF_NAMES):
            setattr(r, name, refs[q].ctypes.data_as(_L.c_double_p))
            setattr(ld, name, 1)
        _L.check("szb_imexop_set_refs", _L.load().szb_imexop_set_refs(self.handle, C.byref(r), C.byref(ld)))
        return self

    def set_refs_device(self, references, stream=None):
        """references: CUDA float64 tensor holding the reference's 42 x Ny column-major `references` block,
        i.e. shape (Ny, 42) C-contiguous (apps/perfect/references.hpp:82-125) -- what a sharded stepper has
        just all-reduced (apps/perfect/perfect.cpp:1397).  No host round trip."""
        assert references.is_cuda and references.shape == (self.n, 42) and references.is_contiguous()
        _L.check("szb_imexop_set_refs_device",
                 _L.load().szb_imexop_set_refs_device(self.handle, _ptr(references), 42, _stream_handle(stream)))
        return self

    def set_isothermal(self, enforce_lower=True, enforce_upper=True, lower=(1.0, 0.0, 0.0, 0.0),
                       upper=(1.0, 0.0, 0.0, 0.0)):
        iso = _L.Isothermal(int(enforce_lower), int(enforce_upper), *map(float, lower), *map(float, upper))
        _L.check("szb_imexop_set_isothermal",
                 _L.load().szb_imexop_set_isothermal(self.handle, C.byref(iso)))
        return self

    def set_nrbc(self, a=None, b=None, c=None):
        arrs = [None if m is None else np.ascontiguousarray(np.asarray(m, dtype=np.float64).reshape(-1))
                for m in (a, b, c)]
        ptrs = [None if m is None else m.ctypes.data_as(_L.c_double_p) for m in arrs]
        _L.check("szb_imexop_set_nrbc", _L.load().szb_imexop_set_nrbc(self.handle, *ptrs))
        return self

    def set_linearization(self, mode="rhome_xyz"):
        """linearize::rhome_xyz (default) or linearize::rhome_y, the wavenumber-independent operator
        (apps/perfect/operator_hybrid_isothermal.cpp:691-761)."""
        code = {"rhome_xyz": 0, "rhome_y": 1}[mode]
        _L.check("szb_imexop_set_linearization", _L.load().szb_imexop_set_linearization(self.handle, code))
        return self

    # ---- batched, device tensors ------------------------------------------------
    def accumulate_batch(self, phi, km, kn, x, beta, y, index=None, x_strides=None,
                         y_strides=None, stream=None):
        """y <- (M + phi L) x + beta y on device tensors.
        x, y: complex128 CUDA tensors; default layout (npencil, 5, n) contiguous,
        i.e. field stride n and pencil stride 5n (interleaved-state pencils)."""
        n = self.n
        xs = x_strides or (n, 5 * n)
        ys = y_strides or (n, 5 * n)
        rc = _L.load().szb_imexop_accumulate_batch(
            self.handle, _d2(phi), int(km.numel()), _ptr(km), _ptr(kn), _ptr(index),
            _ptr(x), xs[0], xs[1], _d2(beta), _ptr(y), ys[0], ys[1], _stream_handle(stream))
        _L.check("szb_imexop_accumulate_batch", rc)
        return y

    def pack_batch(self, phi, km, kn, out, packf=False, with_bc=False, stream=None):
        rc = _L.load().szb_imexop_pack_batch(self.handle, _d2(phi), int(km.numel()), _ptr(km),
                                             _ptr(kn), int(packf), int(with_bc), _ptr(out),
                                             _stream_handle(stream))
        _L.check("szb_imexop_pack_batch", rc)
        return out

    def invert_batch(self, spec: SolverSpec, phi, km, kn, state, index=None, strides=None,
                     extra=None, ipiv=None, info=None, iters=None, stream=None):
        n = self.n
        st = strides or (n, 5 * n)
        cs = spec.c()
        nextra = 0 if extra is None else int(extra.shape[1])
        rc = _L.load().szb_imexop_invert_batch(
            self.handle, C.byref(cs), _d2(phi), int(km.numel()), _ptr(km), _ptr(kn), _ptr(index),
            _ptr(state), st[0], st[1], nextra, _ptr(extra), _ptr(ipiv), _ptr(info), _ptr(iters),
            _stream_handle(stream))
        _L.check("szb_imexop_invert_batch", rc)
        return state

    def workspace_bytes(self):
        return int(_L.load().szb_imexop_workspace_bytes(self.handle))


def wavegrid(Nx, Nz, Lx, Lz, dealias=1.5, xrange=None, zrange=None):
    """specification_grid / pencil_grid extents for one rank owning
    [xrange) x [zrange) of wave space (defaults: everything).  dN = dealias*N,
    wave-space x extent is dNx/2+1 (real-to-complex transform)."""
    dNx, dNz = int(Nx * dealias), int(Nz * dealias)
    xb, xe = xrange if xrange is not None else (0, dNx // 2 + 1)
    zb, ze = zrange if zrange is not None else (0, dNz)
    return _L.WaveGrid(Nx, dNx, xb, xe, Nz, dNz, zb, ze, float(Lx), float(Lz))


def wavenumbers(g):
    L = _L.load()
    npen = L.szb_wavegrid_npencils(C.byref(g))
    km, kn = np.empty(npen), np.empty(npen)
    act = np.empty(npen, dtype=np.int32)
    L.szb_wavegrid_wavenumbers(C.byref(g), km.ctypes.data_as(_L.c_double_p),
                               kn.ctypes.data_as(_L.c_double_p), act.ctypes.data_as(_L.c_int_p))
    return km, kn, act.astype(bool)


def collect_references(scenario, beta, sphys, Ny=None, y0=0, chi=None, top_is_inviscid=False, group=None,
                       out=None, stream=None):
    """Reference profiles from the physical-space state: collect_references (apps/perfect/perfect.cpp:1266-1400).
    sphys: device tensor (5, ny, nz, nx) float64, fields e, mx, my, mz, rho on this rank's planes [y0, y0 + ny) of
    Ny; scenario: dict with Ma, alpha, gamma (Re, Pr unused); beta: viscosity exponent; chi = 1 / (dNx dNz), default
    from the local (nz, nx) (right for a single rank or a y-decomposition).  With torch.distributed initialised (or
    `group` given) the sums are all-reduced over ranks before the scaling, as the reference's MPI_Allreduce.
    Returns the (Ny, 42) device tensor whose memory is the reference's 42 x Ny column-major block: feed it to
    ImexOp.set_refs_device."""
    import torch
    import torch.distributed as dist
    assert sphys.is_cuda and sphys.dtype == torch.float64 and sphys.dim() == 4 and sphys.shape[0] == 5
    assert sphys[0].is_contiguous()
    ny, nz, nx = (int(v) for v in sphys.shape[1:])
    Ny = ny if Ny is None else int(Ny)
    chi = 1.0 / (nz * nx) if chi is None else float(chi)
    multi = group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    s = _L.Scenario(float(scenario.get("Re", 0.0)), float(scenario.get("Pr", 0.0)), float(scenario["Ma"]),
                    float(scenario["alpha"]), float(scenario["gamma"]))
    st = stream or torch.cuda.current_stream(sphys.device)
    L = _L.load()
    need = C.c_size_t(0)
    _L.check("szb_collect_references_device",
             L.szb_collect_references_device(C.byref(s), float(beta), Ny, int(y0), ny, nz * nx, None, 0, 0, 1.0, None, None,
                                             0, C.byref(need), None))
    work = torch.empty(max(need.value, 8) // 8, dtype=torch.float64, device=sphys.device)
    refs = out if out is not None else torch.empty((Ny, 42), dtype=torch.float64, device=sphys.device)
    assert refs.is_contiguous() and refs.shape == (Ny, 42)
    rc = L.szb_collect_references_device(C.byref(s), float(beta), Ny, int(y0), ny, nz * nx, C.c_void_p(sphys.data_ptr()),
                                         sphys.stride(0), int(bool(top_is_inviscid)), 1.0 if multi else chi,
                                         C.c_void_p(refs.data_ptr()), C.c_void_p(work.data_ptr()), work.numel() * 8,
                                         C.byref(need), C.c_void_p(st.cuda_stream))
    _L.check("szb_collect_references_device", rc)
    if multi:
        dist.all_reduce(refs, op=dist.ReduceOp.SUM, group=group)
        refs.mul_(chi)
    return refs


def bsplineop_accumulate_complex_batch(bop, d, alpha, x, beta, y, stream=None):
    """y <- alpha D^(d) x + beta y over the rows of the device tensors x, y (nrhs, n) complex128:
    suzerain_bsplineop_accumulate_complex as batched by operator_tools.hpp:77-116."""
    import torch
    assert x.is_cuda and y.is_cuda and x.dtype == torch.complex128 and y.dtype == torch.complex128
    nrhs = x.shape[0]
    st = stream or torch.cuda.current_stream(x.device)
    a2 = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    b2 = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
    rc = _L.load().szb_bsplineop_accumulate_complex_batch(
        bop.handle, int(d), int(nrhs), a2, C.c_void_p(x.data_ptr()), x.stride(0), b2,
        C.c_void_p(y.data_ptr()), y.stride(0), C.c_void_p(st.cuda_stream))
    _L.check("szb_bsplineop_accumulate_complex_batch", rc)
    return y


def bsplineop_accumulate_batch(bop, d, alpha, x, beta, y, stream=None):
    """Real pencils: y <- alpha D^(d) x + beta y over the rows of the float64 device tensors x, y (nrhs, n):
    suzerain_bsplineop_accumulate (suzerain/bsplineop.c:222-258)."""
    import torch
    assert x.is_cuda and y.is_cuda and x.dtype == torch.float64 and y.dtype == torch.float64
    st = stream or torch.cuda.current_stream(x.device)
    rc = _L.load().szb_bsplineop_accumulate_batch(
        bop.handle, int(d), int(x.shape[0]), float(alpha), C.c_void_p(x.data_ptr()), x.stride(0), float(beta),
        C.c_void_p(y.data_ptr()), y.stride(0), C.c_void_p(st.cuda_stream))
    _L.check("szb_bsplineop_accumulate_batch", rc)
    return y


def bsplineop_apply_batch(bop, d, alpha, x, stream=None):
    """In place: x <- alpha D^(d) x over the rows of the device tensor x (nrhs, n), float64 or complex128:
    suzerain_bsplineop_apply / suzerain_bsplineop_apply_complex (suzerain/bsplineop.c:299-381)."""
    import torch
    assert x.is_cuda and x.dtype in (torch.float64, torch.complex128)
    st = stream or torch.cuda.current_stream(x.device)
    name = "szb_bsplineop_apply_complex_batch" if x.dtype == torch.complex128 else "szb_bsplineop_apply_batch"
    rc = getattr(_L.load(), name)(bop.handle, int(d), int(x.shape[0]), float(alpha), C.c_void_p(x.data_ptr()),
                                  x.stride(0), C.c_void_p(st.cuda_stream))
    _L.check(name, rc)
    return x


def diffwave_apply(dxcnt, dzcnt, alpha, x, grid, stream=None):
    """x <- alpha (i kx)^dxcnt (i kz)^dzcnt x in place on the device tensor x (nz, nx, Ny):
    suzerain_diffwave_apply (suzerain/diffwave.c:65-129)."""
    import torch
    st = stream or torch.cuda.current_stream(x.device)
    a2 = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    rc = _L.load().szb_diffwave_apply_batch(int(dxcnt), int(dzcnt), a2, C.c_void_p(x.data_ptr()),
                                            C.byref(grid), int(x.shape[-1]), C.c_void_p(st.cuda_stream))
    _L.check("szb_diffwave_apply_batch", rc)
    return x


def diffwave_accumulate(dxcnt, dzcnt, alpha, x, beta, y, grid, stream=None):
    """y <- alpha (i kx)^dxcnt (i kz)^dzcnt x + beta y on device tensors (nz, nx, Ny):
    suzerain_diffwave_accumulate (suzerain/diffwave.c:131-198)."""
    import torch
    st = stream or torch.cuda.current_stream(x.device)
    a2 = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    b2 = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
    rc = _L.load().szb_diffwave_accumulate_batch(int(dxcnt), int(dzcnt), a2, C.c_void_p(x.data_ptr()), b2,
                                                 C.c_void_p(y.data_ptr()), C.byref(grid), int(x.shape[-1]),
                                                 C.c_void_p(st.cuda_stream))
    _L.check("szb_diffwave_accumulate_batch", rc)
    return y


class OperatorHybridIsothermal:
    """The three virtuals of operator_hybrid_isothermal on HOST state arrays
    (numpy, optionally pinned); copies to the device and back inside each call
    (apps/perfect/operator_hybrid_isothermal.cpp:103,243,528)."""

    def __init__(self, imexop: ImexOp, grid, spec: SolverSpec | None = None):
        self.op = imexop
        self.grid = grid
        self.spec = spec or SolverSpec()

    def apply_mass_plus_scaled_operator(self, phi, state):
        rc = _L.load().szb_operator_apply_mass_plus_scaled_operator(
            self.op.handle, C.byref(self.grid), _d2(phi), _ptr(state))
        _L.check("szb_operator_apply_mass_plus_scaled_operator", rc)
        return state

    def accumulate_mass_plus_scaled_operator(self, phi, input, beta, output, out_field_stride):
        rc = _L.load().szb_operator_accumulate_mass_plus_scaled_operator(
            self.op.handle, C.byref(self.grid), _d2(phi), _ptr(input), _d2(beta), _ptr(output),
            int(out_field_stride))
        _L.check("szb_operator_accumulate_mass_plus_scaled_operator", rc)
        return output

    def invert_mass_plus_scaled_operator(self, phi, state, ic0=None):
        bad = C.c_int(-1)
        cs = self.spec.c()
        nc = 0 if ic0 is None else int(ic0.shape[0])
        rc = _L.load().szb_operator_invert_mass_plus_scaled_operator(
            self.op.handle, C.byref(cs), C.byref(self.grid), _d2(phi), _ptr(state), nc, _ptr(ic0),
            C.byref(bad))
        if rc > 0:
            # bsmbsm_solver.cpp:123-141: decode the singular row
            N, n = self.op.N, self.op.n
            row = rc - 1
            q = (row % 5) * n + row // 5
            raise _L.SzbError(
                f"invert: pencil {bad.value}: singularity in PAP^T row {row} corresponding to "
                f"A row {q} for state scalar {q // n}", rc)
        _L.check("szb_operator_invert_mass_plus_scaled_operator", rc)
        return state


class OperatorHybridIsothermalDevice:
    """operator_hybrid_isothermal for a DEVICE-resident state (torch CUDA
    complex128 tensors in the reference's interleaved layout
    [5][Ny][Nx_loc][Nz_loc], i.e. shape (npencil, 5, Ny) C-contiguous).  The
    wavenumber tables and the active / dealiased pencil lists are built once
    (operator_hybrid_isothermal.cpp:120-134,163-170) and kept on the device."""

    def __init__(self, imexop: ImexOp, grid, spec: SolverSpec | None = None, device=None):
        import torch
        self.op, self.grid = imexop, grid
        self.spec = spec or SolverSpec()
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        km, kn, act = wavenumbers(grid)
        self.npencil = len(km)
        idx = np.arange(self.npencil, dtype=np.int32)
        self.h_active, self.h_inactive = idx[act], idx[~act]
        self.h_km, self.h_kn = km[act], kn[act]
        self.km = torch.from_numpy(self.h_km).to(self.device)
        self.kn = torch.from_numpy(self.h_kn).to(self.device)
        self.active = torch.from_numpy(self.h_active).to(self.device)
        self.inactive = torch.from_numpy(self.h_inactive).to(self.device)
        self.nactive = int(act.sum())
        self.info = torch.zeros(max(self.nactive, 1), dtype=torch.int32, device=self.device)

    def apply_mass_plus_scaled_operator(self, phi, state, stream=None):
        return self.op.accumulate_batch(phi, self.km, self.kn, state, 0.0, state,
                                        index=self.active, stream=stream)

    def accumulate_mass_plus_scaled_operator(self, phi, input, beta, output, stream=None, interleaved_output=False):
        """output: contiguous state (5, npencil, Ny) (suzerain/storage.hpp:267-268) as in the reference, or
        -- for a device-resident stepper that swaps the two buffers -- interleaved like the input."""
        n = self.op.n
        return self.op.accumulate_batch(phi, self.km, self.kn, input, beta, output, index=self.active,
                                        y_strides=None if interleaved_output else (n * self.npencil, n),
                                        stream=stream)

    def raise_on_singular(self):
        """The device-resident invert records zgbtrf's info per active pencil in ``self.info`` without
        stopping; call this (it synchronises) where a singular operator must be fatal, as it is in the
        reference (bsmbsm_solver.cpp:123-141 -> SUZERAIN_ERROR).  lowstorage.step does so once per step."""
        info = self.info[:self.nactive]
        if self.nactive and bool((info != 0).any()):
            import torch
            bad = int(torch.nonzero(info)[0])
            rc = int(info[bad])
            n = self.op.n
            row = rc - 1
            q = (row % 5) * n + row // 5
            raise _L.SzbError(
                f"invert: pencil {int(self.h_active[bad])}: singularity in PAP^T row {row} corresponding to "
                f"A row {q} for state scalar {q // n}", rc)

    def exchange(self, a, b, stream=None):
        """a <-> b between the interleaved state a (npencil, 5, Ny) and the contiguous state b (5, npencil, Ny):
        `b.exchange(a)` of lowstorage::step (suzerain/lowstorage.hpp:1511), every stored pencil."""
        n = self.op.n
        rc = _L.load().szb_state_exchange(self.npencil, None, 5, n, _ptr(a), n, 5 * n,
                                          _ptr(b), n * self.npencil, n, _stream_handle(stream))
        _L.check("szb_state_exchange", rc)

    def invert_mass_plus_scaled_operator(self, phi, state, stream=None, ipiv=None, iters=None):
        n = self.op.n
        if len(self.h_inactive):
            rc = _L.load().szb_zero_pencils(len(self.h_inactive), _ptr(self.inactive), 5, n,
                                            _ptr(state), n, 5 * n, _stream_handle(stream))
            _L.check("szb_zero_pencils", rc)
        return self.op.invert_batch(self.spec, phi, self.km, self.kn, state, index=self.active,
                                    info=self.info, ipiv=ipiv, iters=iters, stream=stream)
