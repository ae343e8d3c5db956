"""ctypes bindings onto oracle/_ref/libsuzerain_ref.so.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The library is the
reference's own hot-path C sources, compiled unmodified by oracle/Makefile,
plus oracle/ref_shim/ref_glue.c.  Structures below mirror
suzerain/rholut_imexop.h:66-133 and suzerain/bsplineop.h:125-180 field for
field so that the reference functions can be called directly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsuzerain_ref.so")

REF_NAMES = (
    "ux uy uz u2 uxux uxuy uxuz uyuy uyuz uzuz nu nuux nuuy nuuz nuu2 "
    "nuuxux nuuxuy nuuxuz nuuyuy nuuyuz nuuzuz ex_gradrho ey_gradrho "
    "ez_gradrho e_divm e_deltarho").split()
assert len(REF_NAMES) == 26

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class Scenario(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("Re", "Pr", "Ma", "alpha", "gamma")]


class Ref(C.Structure):
    _fields_ = [(n, c_double_p) for n in REF_NAMES]


class RefLd(C.Structure):
    _fields_ = [(n, C.c_int) for n in REF_NAMES]


class Workspace(C.Structure):
    _fields_ = [("method", C.c_int), ("k", C.c_int), ("n", C.c_int),
                ("nderiv", C.c_int), ("kl", c_int_p), ("ku", c_int_p),
                ("max_kl", C.c_int), ("max_ku", C.c_int), ("ld", C.c_int),
                ("D_T", C.POINTER(c_double_p))]


class Bc(C.Structure):
    _fields_ = [("enforce_lower", C.c_int), ("enforce_upper", C.c_int),
                ("E_factor", C.c_double * 2), ("vel_factor", (C.c_double * 3) * 2)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        _lib.ref_set_blas_threads(1)
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Problem:
    """Keeps numpy storage alive behind the reference's structs."""

    def __init__(self, op, scenario: dict, refs: np.ndarray, bc: dict | None = None,
                 nrbc=None):
        """op: oracle.bspline.BsplineOp; refs: (26, n) float64 in REF_NAMES order.
        bc: dict(enforce_lower, enforce_upper, E_factor[2], vel_factor[2][3]).
        nrbc: None or (a, b, c) each 5x5 column-major float64."""
        self.op = op
        self.n = op.n
        self.scen = Scenario(**{k: float(scenario[k]) for k in ("Re", "Pr", "Ma", "alpha", "gamma")})
        self.refs = np.ascontiguousarray(refs, dtype=np.float64)
        assert self.refs.shape == (26, op.n)
        self.ref = Ref()
        self.refld = RefLd()
        for q, name in enumerate(REF_NAMES):
            setattr(self.ref, name, _p(self.refs[q]))
            setattr(self.refld, name, 1)
        # workspace: only D_T[0..2] are read by the hot path (rholut_imexop.c:77)
        nd = op.nderiv
        self._storage = np.ascontiguousarray(op.storage)
        self._kl = np.ascontiguousarray(op.kl, dtype=np.int32)
        self._ku = np.ascontiguousarray(op.ku, dtype=np.int32)
        self._DT = (c_double_p * (nd + 1))()
        base = self._storage.ctypes.data
        for d in range(nd + 1):
            addr = base + 8 * (d * op.n * op.ld + op.D_T_offset(d))
            self._DT[d] = C.cast(addr, c_double_p)
        self.w = Workspace(0, op.k, op.n, nd, _p(self._kl, C.c_int), _p(self._ku, C.c_int),
                           op.max_kl, op.max_ku, op.ld, self._DT)
        self.bc = None
        if bc is not None:
            self.bc = Bc()
            self.bc.enforce_lower = int(bc.get("enforce_lower", 1))
            self.bc.enforce_upper = int(bc.get("enforce_upper", 1))
            for i in range(2):
                self.bc.E_factor[i] = float(bc["E_factor"][i])
                for j in range(3):
                    self.bc.vel_factor[i][j] = float(bc["vel_factor"][i][j])
        self.nrbc = None
        if nrbc is not None:
            self.nrbc = tuple(None if m is None else np.ascontiguousarray(
                np.asarray(m, dtype=np.float64).reshape(-1)) for m in nrbc)
        self.S = 5
        self.N = 5 * op.n
        self.KL = 5 * (op.max_kl + 1) - 1
        self.KU = 5 * (op.max_ku + 1) - 1
        self.LD = self.KL + 1 + self.KU

    def _abc(self):
        if self.nrbc is None:
            return None, None, None
        return tuple(_p(m) for m in self.nrbc)

    def assemble(self, phi: complex, km: float, kn: float, packf=False, with_bc=True):
        rows = self.LD + (self.KL if packf else 0)
        out = np.full((self.N, rows), np.nan + 1j * np.nan, dtype=np.complex128)
        phi2 = (C.c_double * 2)(phi.real, phi.imag)
        a, b, c = self._abc()
        rc = lib().ref_assemble(phi2, C.c_double(km), C.c_double(kn),
                                C.byref(self.scen), C.byref(self.ref), C.byref(self.refld),
                                C.byref(self.w),
                                C.byref(self.bc) if (with_bc and self.bc is not None) else None,
                                a, b, c, int(packf), out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return out                      # [column j, band row]

    def invert(self, solver: str, phi: complex, km, kn, state, extra=None, nthreads=1,
               want_ipiv=False, want_iters=False):
        """state: (npencil, 5*n) complex128, solved in place (copy returned)."""
        km = np.ascontiguousarray(km, dtype=np.float64)
        kn = np.ascontiguousarray(kn, dtype=np.float64)
        x = np.array(state, dtype=np.complex128, order="C", copy=True)
        npencil = x.shape[0]
        assert x.shape == (npencil, self.N) and km.shape == (npencil,) == kn.shape
        nextra = 0
        ex = None
        if extra is not None:
            ex = np.array(extra, dtype=np.complex128, order="C", copy=True)
            nextra = ex.shape[1]
            assert ex.shape == (npencil, nextra, self.N)
        ipiv = np.zeros((npencil, self.N), dtype=np.int32) if want_ipiv else None
        iters = np.zeros(npencil, dtype=np.int32) if want_iters else None
        bad = C.c_int(-1)
        phi2 = (C.c_double * 2)(phi.real, phi.imag)
        a, b, c = self._abc()
        assert self.bc is not None
        info = lib().ref_invert_batch(
            {"zgbsv": 0, "zcgbsvx": 1}[solver], phi2,
            C.byref(self.scen), C.byref(self.ref), C.byref(self.refld), C.byref(self.w),
            C.byref(self.bc), a, b, c, npencil, _p(km), _p(kn),
            x.ctypes.data_as(C.c_void_p), nextra,
            ex.ctypes.data_as(C.c_void_p) if ex is not None else None,
            _p(ipiv, C.c_int), _p(iters, C.c_int), int(nthreads), C.byref(bad))
        res = {"x": x, "info": info, "first_bad": bad.value}
        if ex is not None:
            res["extra"] = ex
        if want_ipiv:
            res["ipiv"] = ipiv
        if want_iters:
            res["iters"] = iters
        return res

    def invert_spec(self, phi: complex, km, kn, state, method="zcgbsvx", equil=False, reuse=False,
                    aiter=1, siter=-1, diter=5, tolsc=0.0, rowlen=1, nthreads=1):
        """Any solver specification of apps/perfect/test_implicit_solvers.sh:24-30, with the
        solver state carried along rows of ``rowlen`` consecutive pencils as the reference's
        kx loop does (so that reuse=true takes its approximate-factorisation path)."""
        km = np.ascontiguousarray(km, dtype=np.float64)
        kn = np.ascontiguousarray(kn, dtype=np.float64)
        x = np.array(state, dtype=np.complex128, order="C", copy=True)
        npencil = x.shape[0]
        assert x.shape == (npencil, self.N) and km.shape == (npencil,) == kn.shape
        iters = np.zeros((npencil, 2), dtype=np.int32)
        stats = np.zeros((npencil, 2), dtype=np.float64)
        bad = C.c_int(-1)
        phi2 = (C.c_double * 2)(complex(phi).real, complex(phi).imag)
        a, b, c = self._abc()
        assert self.bc is not None
        f = lib().ref_invert_spec_batch
        f.restype = C.c_int
        info = f(C.c_int({"zgbsv": 0, "zcgbsvx": 1, "zgbsvx": 2}[method]), C.c_int(int(equil)), C.c_int(int(reuse)),
                 C.c_int(aiter), C.c_int(siter), C.c_int(diter), C.c_double(tolsc), phi2,
                 C.byref(self.scen), C.byref(self.ref), C.byref(self.refld), C.byref(self.w),
                 C.byref(self.bc), a, b, c, C.c_int(npencil), C.c_int(rowlen), _p(km), _p(kn),
                 x.ctypes.data_as(C.c_void_p), _p(iters, C.c_int), _p(stats), C.c_int(int(nthreads)), C.byref(bad))
        return {"x": x, "info": info, "first_bad": bad.value, "iters": iters, "stats": stats}

    def accumulate(self, phi: complex, km, kn, x, beta: complex = 0.0, y=None, nthreads=1):
        km = np.ascontiguousarray(km, dtype=np.float64)
        kn = np.ascontiguousarray(kn, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.complex128)
        npencil = x.shape[0]
        out = (np.zeros_like(x) if y is None
               else np.array(y, dtype=np.complex128, order="C", copy=True))
        phi2 = (C.c_double * 2)(complex(phi).real, complex(phi).imag)
        beta2 = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
        a, b, c = self._abc()
        lib().ref_accumulate_batch(phi2, C.byref(self.scen), C.byref(self.ref),
                                   C.byref(self.refld), C.byref(self.w), a, b, c,
                                   npencil, _p(km), _p(kn),
                                   x.ctypes.data_as(C.c_void_p), beta2,
                                   out.ctypes.data_as(C.c_void_p), int(nthreads))
        return out

    # ---- linearize::rhome_y: the wavenumber-independent "00" operator ----
    def assemble00(self, phi: complex, packf=False, with_bc=True):
        rows = self.LD + (self.KL if packf else 0)
        out = np.full((self.N, rows), np.nan + 1j * np.nan, dtype=np.complex128)
        phi2 = (C.c_double * 2)(complex(phi).real, complex(phi).imag)
        _, _, c = self._abc()
        rc = lib().ref_assemble00(phi2, C.byref(self.scen), C.byref(self.ref), C.byref(self.refld),
                                  C.byref(self.w),
                                  C.byref(self.bc) if (with_bc and self.bc is not None) else None, c,
                                  int(packf), out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return out

    def invert00(self, phi: complex, state, want_ipiv=False):
        """operator_hybrid_isothermal.cpp:691-761: one factorisation of the 00 operator, one
        zgbtrs('T') per pencil."""
        x = np.array(state, dtype=np.complex128, order="C", copy=True)
        npencil = x.shape[0]
        N = x.shape[1]
        phi2 = (C.c_double * 2)(complex(phi).real, complex(phi).imag)
        ipiv = np.zeros(N, dtype=np.int32)
        _, _, c = self._abc()
        info = lib().ref_invert00_batch(phi2, C.byref(self.scen), C.byref(self.ref), C.byref(self.refld),
                                        C.byref(self.w), C.byref(self.bc), c, npencil,
                                        x.ctypes.data_as(C.c_void_p), _p(ipiv, C.c_int))
        res = dict(x=x, info=int(info))
        if want_ipiv:
            res["ipiv"] = ipiv
        return res

    def accumulate00(self, phi: complex, x, beta: complex = 0.0, y=None):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        out = (np.zeros_like(x) if y is None else np.array(y, dtype=np.complex128, order="C", copy=True))
        phi2 = (C.c_double * 2)(complex(phi).real, complex(phi).imag)
        beta2 = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
        _, _, c = self._abc()
        lib().ref_accumulate00_batch(phi2, C.byref(self.scen), C.byref(self.ref), C.byref(self.refld),
                                     C.byref(self.w), c, x.shape[0], x.ctypes.data_as(C.c_void_p), beta2,
                                     out.ctypes.data_as(C.c_void_p))
        return out

    def bsplineop_accumulate_complex(self, d, alpha: complex, x, beta: complex = 0.0, y=None):
        """y <- alpha D^(d) x + beta y per row of x (nrhs, n): suzerain_bsplineop_accumulate_complex
        (suzerain/bsplineop.c:260-297) through the reference's own zgbmv_d_z."""
        x = np.ascontiguousarray(x, dtype=np.complex128)
        nrhs, n = x.shape
        out = (np.zeros_like(x) if y is None else np.array(y, dtype=np.complex128, order="C", copy=True))
        a2 = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
        b2 = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
        lib().ref_bsplineop_accumulate_complex(int(d), int(nrhs), a2, x.ctypes.data_as(C.c_void_p), 1, n, b2,
                                               out.ctypes.data_as(C.c_void_p), 1, n, C.byref(self.w))
        return out

    def bsplineop_accumulate(self, d, alpha: float, x, beta: float = 0.0, y=None):
        """Real coefficients: suzerain_bsplineop_accumulate (suzerain/bsplineop.c:222-258)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        nrhs, n = x.shape
        out = (np.zeros_like(x) if y is None else np.array(y, dtype=np.float64, order="C", copy=True))
        lib().ref_bsplineop_accumulate(int(d), int(nrhs), C.c_double(alpha), x.ctypes.data_as(C.c_void_p), 1, n,
                                       C.c_double(beta), out.ctypes.data_as(C.c_void_p), 1, n, C.byref(self.w))
        return out

    def bsplineop_apply(self, d, alpha: float, x):
        """In place, real or complex rows: suzerain_bsplineop_apply / _apply_complex (suzerain/bsplineop.c:299-381)."""
        cplx = np.iscomplexobj(x)
        out = np.array(x, dtype=np.complex128 if cplx else np.float64, order="C", copy=True)
        nrhs, n = out.shape
        f = lib().ref_bsplineop_apply_complex if cplx else lib().ref_bsplineop_apply
        f(int(d), int(nrhs), C.c_double(alpha), out.ctypes.data_as(C.c_void_p), 1, n, C.byref(self.w))
        return out


def zgbsv_T(N, KL, KU, LU, B):
    """In-place zgbtrf + zgbtrs('T') on LAPACK band storage.
    LU: (N, 2KL+KU+1) complex128 C-contiguous == column-major (2KL+KU+1) x N.
    B: (nrhs, N).  Returns (LU, ipiv, X, info)."""
    LU = np.array(LU, dtype=np.complex128, order="C", copy=True)
    B = np.array(B, dtype=np.complex128, order="C", copy=True).reshape(-1, N)
    ipiv = np.zeros(N, dtype=np.int32)
    info = lib().ref_zgbsv_T(N, KL, KU, LU.ctypes.data_as(C.c_void_p), 2 * KL + KU + 1,
                             _p(ipiv, C.c_int), B.ctypes.data_as(C.c_void_p), B.shape[0])
    return LU, ipiv, B, info


def q(S, n, i):
    return lib().suzerain_bsmbsm_q(S, n, i) if hasattr(lib(), "suzerain_bsmbsm_q") else (i % S) * n + i // S


def diffwave(dxcnt, dzcnt, alpha: complex, x, Lx, Lz, grid, beta: complex | None = None, y=None):
    """suzerain_diffwave_apply (beta is None: returns alpha D x) or suzerain_diffwave_accumulate
    (returns alpha D x + beta y), suzerain/diffwave.c:65-198.  x, y: (nz, nx, Ny) complex128 with
    grid = (Nx, dNx, dkbx, dkex, Nz, dNz, dkbz, dkez)."""
    L = lib()
    Nx, dNx, dkbx, dkex, Nz, dNz, dkbz, dkez = [int(v) for v in grid]
    x = np.ascontiguousarray(x, dtype=np.complex128)
    Ny = x.shape[-1]
    a2 = (C.c_double * 2)(alpha.real, alpha.imag)
    if beta is None:
        out = x.copy()
        L.ref_diffwave_apply(int(dxcnt), int(dzcnt), a2, out.ctypes.data_as(C.c_void_p), C.c_double(Lx),
                             C.c_double(Lz), Ny, Nx, dNx, dkbx, dkex, Nz, dNz, dkbz, dkez)
        return out
    out = np.ascontiguousarray(y, dtype=np.complex128).copy()
    b2 = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
    L.ref_diffwave_accumulate(int(dxcnt), int(dzcnt), a2, x.ctypes.data_as(C.c_void_p), b2,
                              out.ctypes.data_as(C.c_void_p), C.c_double(Lx), C.c_double(Lz), Ny,
                              Nx, dNx, dkbx, dkex, Nz, dNz, dkbz, dkez)
    return out
