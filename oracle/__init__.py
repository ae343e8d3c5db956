"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the B200 implicit-operator path:

* ``oracle.ref``     ctypes bindings onto ``oracle/_ref/libsuzerain_ref.so``, the
                     reference's own unmodified C sources compiled by
                     ``oracle/Makefile`` (LAPACK = OpenBLAS inside SciPy).
* ``oracle.port``    bindings onto ``oracle/liboracle_port.so``, a from-scratch
                     plain-C restatement of the same algorithms
                     (``oracle/oracle_port.c``), pinned against the reference's
                     golden vectors in ``tests/golden`` and against ``oracle.ref``.
* ``oracle.bspline`` SciPy-based construction of the Greville collocation
                     operators (GSL is absent), pinned against the golden
                     matrices of the reference's ``tests/test_bsplineop.cpp``.
* ``oracle.synth``   synthetic channel/boundary-layer inputs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under
``suzerain_b200/`` does.
"""
