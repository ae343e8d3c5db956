"""The drop-in library re-declares the reference's own per-pencil functions: compile its header
against the reference's headers (same names => the C compiler rejects any prototype mismatch), check
the mirrored struct layouts, and check that every symbol is exported.  CPU only; the compile check
needs /root/reference (build container) and is skipped elsewhere."""
import ctypes as C
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SHIM = os.path.join(ROOT, "oracle", "ref_shim", "include")
NAMES = ["suzerain_rholut_imexop_accumulate", "suzerain_rholut_imexop_accumulate00",
         "suzerain_rholut_imexop_packc", "suzerain_rholut_imexop_packc00",
         "suzerain_rholut_imexop_packf", "suzerain_rholut_imexop_packf00"]

SRC_REF = r"""
#include <stddef.h>
#include <suzerain/rholut_imexop.h>          /* the reference's own prototypes and types */
#define SZB_DROPIN_USE_REFERENCE_TYPES
#include "suzerain_b200_dropin.h"            /* re-declares the same functions: must be compatible */

/* the structs the main header mirrors are the reference's, field for field */
_Static_assert(sizeof(szb_rholut_imexop_scenario) == sizeof(suzerain_rholut_imexop_scenario), "scenario");
_Static_assert(offsetof(szb_rholut_imexop_scenario, gamma) == offsetof(suzerain_rholut_imexop_scenario, gamma), "scenario.gamma");
_Static_assert(sizeof(szb_rholut_imexop_ref) == sizeof(suzerain_rholut_imexop_ref), "ref");
_Static_assert(offsetof(szb_rholut_imexop_ref, e_deltarho) == offsetof(suzerain_rholut_imexop_ref, e_deltarho), "ref.e_deltarho");
_Static_assert(offsetof(szb_rholut_imexop_ref, nuuxuy) == offsetof(suzerain_rholut_imexop_ref, nuuxuy), "ref.nuuxuy");
_Static_assert(sizeof(szb_rholut_imexop_refld) == sizeof(suzerain_rholut_imexop_refld), "refld");
_Static_assert(sizeof(szb_bsmbsm) == sizeof(suzerain_bsmbsm), "bsmbsm");
_Static_assert(offsetof(szb_bsmbsm, LD) == offsetof(suzerain_bsmbsm, LD), "bsmbsm.LD");
_Static_assert(offsetof(szb_bsmbsm, KL) == offsetof(suzerain_bsmbsm, KL), "bsmbsm.KL");

/* the workspace struct the drop-in reads in place, as the non-reference build declares it */
struct mirror { int method; int k, n, nderiv; int *kl, *ku; int max_kl, max_ku, ld; double **D_T; };
_Static_assert(sizeof(struct mirror) == sizeof(suzerain_bsplineop_workspace), "workspace size");
_Static_assert(offsetof(struct mirror, k) == offsetof(suzerain_bsplineop_workspace, k), "k");
_Static_assert(offsetof(struct mirror, nderiv) == offsetof(suzerain_bsplineop_workspace, nderiv), "nderiv");
_Static_assert(offsetof(struct mirror, kl) == offsetof(suzerain_bsplineop_workspace, kl), "kl");
_Static_assert(offsetof(struct mirror, ku) == offsetof(suzerain_bsplineop_workspace, ku), "ku");
_Static_assert(offsetof(struct mirror, max_ku) == offsetof(suzerain_bsplineop_workspace, max_ku), "max_ku");
_Static_assert(offsetof(struct mirror, ld) == offsetof(suzerain_bsplineop_workspace, ld), "ld");
_Static_assert(offsetof(struct mirror, D_T) == offsetof(suzerain_bsplineop_workspace, D_T), "D_T");
int main(void) { return 0; }
"""

SRC_PLAIN = r"""
#include "suzerain_b200_dropin.h"
int main(void) { return sizeof(szb_dropin_bsplineop_workspace) > 0 ? 0 : 1; }
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "suzerain")), reason="reference tree not present")
def test_dropin_prototypes_are_the_references():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "abi.c")
        open(src, "w").write(SRC_REF)
        r = subprocess.run(["gcc", "-std=gnu99", "-fsyntax-only", "-Wall", "-Werror", "-I" + SHIM, "-I" + REF,
                            "-I" + os.path.join(ROOT, "include"), src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_dropin_header_is_plain_c_and_cxx():
    with tempfile.TemporaryDirectory() as d:
        for ext, cmd in (("c", ["gcc", "-std=c99"]), ("cpp", ["g++", "-std=c++11"])):
            src = os.path.join(d, "plain." + ext)
            open(src, "w").write(SRC_PLAIN)
            r = subprocess.run(cmd + ["-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), src],
                               capture_output=True, text=True)
            assert r.returncode == 0, r.stderr


def test_dropin_library_exports_the_reference_symbols():
    path = os.path.join(ROOT, "suzerain_b200", "libsuzerain_b200_dropin.so")
    assert os.path.exists(path), "build it with __graft_entry__.build()"
    lib = C.CDLL(path)
    for name in NAMES:
        assert hasattr(lib, name), name
