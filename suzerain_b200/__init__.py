"""suzerain_b200 -- B200-native implicit wall-normal operator path of Suzerain.

The product is ``libsuzerain_b200.so`` (hand-written sm_100a FP64 CUDA behind the
C ABI of ``include/suzerain_b200.h``); this package is the thin host-side mirror
of the reference's operator interface used by tests and ``bench.py``.
"""
from . import lib  # noqa: F401
from . import lowstorage  # noqa: F401
from .pencil import PencilGrid  # noqa: F401
from .api import (BsplineOp, ImexOp, OperatorHybridIsothermal, OperatorHybridIsothermalDevice, SolverSpec,  # noqa: F401
                  bsplineop_accumulate_batch, bsplineop_accumulate_complex_batch, bsplineop_apply_batch,
                  collect_references, diffwave_accumulate, diffwave_apply,
                  htstretch_breakpoints, wavegrid, wavenumbers)

__all__ = ["lib", "lowstorage", "PencilGrid", "BsplineOp", "ImexOp", "OperatorHybridIsothermal", "OperatorHybridIsothermalDevice", "SolverSpec",
           "bsplineop_accumulate_batch", "bsplineop_accumulate_complex_batch", "bsplineop_apply_batch",
           "collect_references", "diffwave_accumulate", "diffwave_apply",
           "htstretch_breakpoints", "wavegrid", "wavenumbers"]
