"""The shim a maintainer adds to apps/perfect (include/suzerain_b200_shim.hpp) and the header-only C++
bsmbsm_solver drop-in (include/suzerain_b200_solver.hpp) compile with -Wall -Werror and link against
libsuzerain_b200.so -- the former against stand-ins for the reference's Boost / Eigen based types
(tests/mock_reference/).  CPU only: nothing is executed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "suzerain_b200")


def _build(src, extra, tmp_path):
    exe = str(tmp_path / (os.path.basename(src) + ".exe"))
    r = subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include")] + extra
                       + [src, "-o", exe, "-L" + LIBDIR, "-lsuzerain_b200", "-Wl,-rpath," + LIBDIR],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_operator_shim_compiles_and_links(tmp_path):
    exe = _build(os.path.join(ROOT, "tests", "cxx", "shim_instantiation.cpp"),
                 ["-I" + os.path.join(ROOT, "tests", "mock_reference")], tmp_path)
    assert subprocess.run([exe]).returncode == 0          # argc < 100: returns before touching a device


def test_solver_protocol_driver_compiles_and_links(tmp_path):
    _build(os.path.join(ROOT, "tests", "cxx", "solver_protocol.cpp"), [], tmp_path)
