"""ctypes loader and prototypes for libsuzerain_b200.so (include/suzerain_b200.h).

The library is the product: hand-written sm_100a CUDA kernels behind a plain C
ABI.  There is no fallback of any kind -- if the shared object is missing or a
CUDA device is unavailable, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsuzerain_b200.so")
# development only: an instrumented / experimental build of the same library (tools/build_variant.sh)
if os.environ.get("SZB_LIB"):
    LIB_PATH = os.path.abspath(os.environ["SZB_LIB"])

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_void_p = C.c_void_p

REF_NAMES = (
    "ux uy uz u2 uxux uxuy uxuz uyuy uyuz uzuz nu nuux nuuy nuuz nuu2 "
    "nuuxux nuuxuy nuuxuz nuuyuy nuuyuz nuuzuz ex_gradrho ey_gradrho "
    "ez_gradrho e_divm e_deltarho").split()


class Bsmbsm(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("S", "n", "kl", "ku", "ld", "N", "KL", "KU", "LD")]


class Scenario(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("Re", "Pr", "Ma", "alpha", "gamma")]


class Ref(C.Structure):
    _fields_ = [(n, c_double_p) for n in REF_NAMES]


class RefLd(C.Structure):
    _fields_ = [(n, C.c_int) for n in REF_NAMES]


class Isothermal(C.Structure):
    _fields_ = [("enforce_lower", C.c_int), ("enforce_upper", C.c_int)] + [
        (n, C.c_double) for n in ("lower_T", "lower_u", "lower_v", "lower_w",
                                  "upper_T", "upper_u", "upper_v", "upper_w")]


class ZgbsvSpec(C.Structure):
    _fields_ = [("method", C.c_int), ("aiter", C.c_int), ("diter", C.c_int),
                ("tolsc", C.c_double), ("equil", C.c_int), ("reuse", C.c_int), ("siter", C.c_int)]


class WaveGrid(C.Structure):
    _fields_ = [("Nx", C.c_int), ("dNx", C.c_int), ("dkbx", C.c_int), ("dkex", C.c_int),
                ("Nz", C.c_int), ("dNz", C.c_int), ("dkbz", C.c_int), ("dkez", C.c_int),
                ("Lx", C.c_double), ("Lz", C.c_double)]


SOLVER_ZGBSV, SOLVER_ZCGBSVX = 0, 1

# name -> (restype, argtypes); every symbol include/suzerain_b200.h declares
D2 = C.c_double * 2
PROTOTYPES = {
    "szb_gbmatrix_offset": (C.c_int, [C.c_int] * 5),
    "szb_gbmatrix_in_band": (C.c_int, [C.c_int] * 4),
    "szb_bsmbsm_construct": (Bsmbsm, [C.c_int] * 4),
    "szb_bsmbsm_q": (C.c_int, [C.c_int] * 3),
    "szb_bsmbsm_qinv": (C.c_int, [C.c_int] * 3),
    "szb_bsmbsm_zaPxpby_batch": (C.c_int, [C.c_char, C.c_int, C.c_int, D2, c_void_p, D2,
                                           c_void_p, C.c_int, c_void_p]),
    "szb_bsplineop_alloc": (C.c_int, [C.c_int, C.c_int, c_double_p, C.c_int, C.POINTER(c_void_p)]),
    "szb_bsplineop_free": (None, [c_void_p]),
    "szb_bsplineop_k": (C.c_int, [c_void_p]),
    "szb_bsplineop_n": (C.c_int, [c_void_p]),
    "szb_bsplineop_nderiv": (C.c_int, [c_void_p]),
    "szb_bsplineop_kl": (C.c_int, [c_void_p, C.c_int]),
    "szb_bsplineop_ku": (C.c_int, [c_void_p, C.c_int]),
    "szb_bsplineop_max_kl": (C.c_int, [c_void_p]),
    "szb_bsplineop_max_ku": (C.c_int, [c_void_p]),
    "szb_bsplineop_ld": (C.c_int, [c_void_p]),
    "szb_bsplineop_D_T": (c_double_p, [c_void_p, C.c_int]),
    "szb_bsplineop_greville": (C.c_int, [c_void_p, c_double_p]),
    "szb_bsplineop_from_storage": (C.c_int, [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p,
                                             c_double_p, C.POINTER(c_void_p)]),
    "szb_htstretch1": (C.c_double, [C.c_double] * 3),
    "szb_htstretch2": (C.c_double, [C.c_double] * 3),
    "szb_bsplineop_accumulate_complex_batch": (C.c_int, [c_void_p, C.c_int, C.c_int, D2, c_void_p,
                                                         C.c_size_t, D2, c_void_p, C.c_size_t,
                                                         c_void_p]),
    "szb_bsplineop_apply_complex_batch": (C.c_int, [c_void_p, C.c_int, C.c_int, C.c_double, c_void_p, C.c_size_t,
                                                    c_void_p]),
    "szb_bsplineop_accumulate_batch": (C.c_int, [c_void_p, C.c_int, C.c_int, C.c_double, c_void_p, C.c_size_t,
                                                 C.c_double, c_void_p, C.c_size_t, c_void_p]),
    "szb_bsplineop_apply_batch": (C.c_int, [c_void_p, C.c_int, C.c_int, C.c_double, c_void_p, C.c_size_t, c_void_p]),
    "szb_collect_references_device": (C.c_int, [c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_size_t, c_void_p,
                                                C.c_size_t, C.c_int, C.c_double, c_void_p, c_void_p, C.c_size_t,
                                                C.POINTER(C.c_size_t), c_void_p]),
    "szb_diffwave_apply_batch": (C.c_int, [C.c_int, C.c_int, D2, c_void_p, c_void_p, C.c_int, c_void_p]),
    "szb_diffwave_accumulate_batch": (C.c_int, [C.c_int, C.c_int, D2, c_void_p, D2, c_void_p, c_void_p,
                                                C.c_int, c_void_p]),
    "szb_zgbsv_spec_default": (ZgbsvSpec, []),
    "szb_imexop_create": (C.c_int, [c_void_p, C.POINTER(c_void_p)]),
    "szb_imexop_destroy": (None, [c_void_p]),
    "szb_imexop_set_scenario": (C.c_int, [c_void_p, C.POINTER(Scenario)]),
    "szb_imexop_set_refs": (C.c_int, [c_void_p, C.POINTER(Ref), C.POINTER(RefLd)]),
    "szb_imexop_set_refs_device": (C.c_int, [c_void_p, c_void_p, C.c_int, c_void_p]),
    "szb_imexop_set_isothermal": (C.c_int, [c_void_p, C.POINTER(Isothermal)]),
    "szb_imexop_set_nrbc": (C.c_int, [c_void_p, c_double_p, c_double_p, c_double_p]),
    "szb_imexop_bsmbsm": (Bsmbsm, [c_void_p]),
    "szb_imexop_set_linearization": (C.c_int, [c_void_p, C.c_int]),
    "szb_imexop_accumulate_batch": (C.c_int, [c_void_p, D2, C.c_int, c_void_p, c_void_p, c_void_p,
                                              c_void_p, C.c_size_t, C.c_size_t, D2,
                                              c_void_p, C.c_size_t, C.c_size_t, c_void_p]),
    "szb_imexop_pack_batch": (C.c_int, [c_void_p, D2, C.c_int, c_void_p, c_void_p, C.c_int,
                                        C.c_int, c_void_p, c_void_p]),
    "szb_imexop_invert_batch": (C.c_int, [c_void_p, C.POINTER(ZgbsvSpec), D2, C.c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, C.c_size_t, C.c_size_t,
                                          C.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "szb_imexop_workspace_bytes": (C.c_size_t, [c_void_p]),
    "szb_bsmbsm_solver_solve": (C.c_int, [C.POINTER(Bsmbsm), C.POINTER(ZgbsvSpec), C.c_char, C.c_int, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "szb_state_exchange": (C.c_int, [C.c_int, c_void_p, C.c_int, C.c_int, c_void_p, C.c_size_t, C.c_size_t,
                                     c_void_p, C.c_size_t, C.c_size_t, c_void_p]),
    "szb_zero_pencils": (C.c_int, [C.c_int, c_void_p, C.c_int, C.c_int, c_void_p, C.c_size_t,
                                   C.c_size_t, c_void_p]),
    "szb_zgbtrf_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, c_void_p, C.c_int, C.c_size_t,
                                   c_void_p, c_void_p, C.c_int, c_void_p]),
    "szb_zgbtrs_batch": (C.c_int, [C.c_char, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p,
                                   C.c_int, C.c_size_t, c_void_p, c_void_p, C.c_int,
                                   C.c_size_t, C.c_int, c_void_p]),
    "szb_zcgbsvx_batch": (C.c_int, [C.c_char, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_double, c_void_p, C.c_size_t, c_void_p, C.c_size_t,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    C.c_int, c_void_p]),
    "szb_rholut_imexop_accumulate": (C.c_int, [D2, C.c_double, C.c_double, C.POINTER(Scenario),
                                               C.POINTER(Ref), C.POINTER(RefLd), c_void_p]
                                     + [c_void_p] * 5 + [D2] + [c_void_p] * 5 + [c_double_p] * 3),
    "szb_rholut_imexop_packc": (C.c_int, [D2, C.c_double, C.c_double, C.POINTER(Scenario),
                                          C.POINTER(Ref), C.POINTER(RefLd), c_void_p,
                                          C.POINTER(Bsmbsm), c_void_p] + [c_double_p] * 3),
    "szb_rholut_imexop_packf": (C.c_int, [D2, C.c_double, C.c_double, C.POINTER(Scenario),
                                          C.POINTER(Ref), C.POINTER(RefLd), c_void_p,
                                          C.POINTER(Bsmbsm), c_void_p] + [c_double_p] * 3),
    "szb_rholut_imexop_accumulate00": (C.c_int, [D2, C.POINTER(Scenario), C.POINTER(Ref), C.POINTER(RefLd), c_void_p]
                                       + [c_void_p] * 5 + [D2] + [c_void_p] * 5 + [c_double_p]),
    "szb_rholut_imexop_packc00": (C.c_int, [D2, C.POINTER(Scenario), C.POINTER(Ref), C.POINTER(RefLd), c_void_p,
                                            C.POINTER(Bsmbsm), c_void_p, c_double_p]),
    "szb_rholut_imexop_packf00": (C.c_int, [D2, C.POINTER(Scenario), C.POINTER(Ref), C.POINTER(RefLd), c_void_p,
                                            C.POINTER(Bsmbsm), c_void_p, c_double_p]),
    "szb_wavegrid_npencils": (C.c_int, [C.POINTER(WaveGrid)]),
    "szb_wavegrid_nactive": (C.c_int, [C.POINTER(WaveGrid)]),
    "szb_wavegrid_wavenumbers": (C.c_int, [C.POINTER(WaveGrid), c_double_p, c_double_p, c_int_p]),
    "szb_operator_apply_mass_plus_scaled_operator": (C.c_int, [c_void_p, C.POINTER(WaveGrid), D2,
                                                               c_void_p]),
    "szb_operator_accumulate_mass_plus_scaled_operator": (C.c_int, [c_void_p, C.POINTER(WaveGrid),
                                                                    D2, c_void_p, D2, c_void_p,
                                                                    C.c_size_t]),
    "szb_operator_invert_mass_plus_scaled_operator": (C.c_int, [c_void_p, C.POINTER(ZgbsvSpec),
                                                                C.POINTER(WaveGrid), D2, c_void_p,
                                                                C.c_int, c_void_p, c_int_p]),
    "szb_device_count": (C.c_int, []),
    "szb_version": (C.c_char_p, []),
    "szb_launch_count": (C.c_ulonglong, []),
}

_lib = None


def load():
    """Load the C-ABI library; raises OSError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (or `make -C suzerain_b200/csrc`).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if a declared symbol is absent
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class SzbError(RuntimeError):
    def __init__(self, fn, code):
        self.code = code
        if code <= -1000:
            msg = f"{fn}: CUDA runtime error {-(code + 1000)}"
        elif code < 0:
            msg = f"{fn}: invalid argument {-code}"
        else:
            msg = f"{fn}: numerical failure, info={code}"
        super().__init__(msg)


def check(fn, code):
    if code != 0:
        raise SzbError(fn, code)
