#!/usr/bin/env python
"""Small driver for ncu captures: one invert (+ accumulate) launch over a bounded
number of pencils of a named grid.  python tools/prof_invert.py [config] [npencils] [solver]
(SZB_LINEARIZATION=rhome_y selects the wavenumber-independent linearisation.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import parity_common as pc

cfg = sys.argv[1] if len(sys.argv) > 1 else "channel_192x96x192"
npen = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
solver = sys.argv[3] if len(sys.argv) > 3 else "zgbsv"
case = pc.make_case(cfg, max_pencils=npen)
dev = torch.device("cuda:0")
for rep in range(2):
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import suzerain_b200 as sz
    op = pc.make_imexop(case)
    op.set_linearization(os.environ.get("SZB_LINEARIZATION", "rhome_xyz"))
    km = torch.from_numpy(case.km).to(dev); kn = torch.from_numpy(case.kn).to(dev)
    st = torch.from_numpy(case.x.copy()).to(dev)
    info = torch.zeros(len(case.km), dtype=torch.int32, device=dev)
    op.invert_batch(sz.SolverSpec(method=solver), case.phi, km, kn, st, info=info)   # warm / alloc
    t0.record()
    op.invert_batch(sz.SolverSpec(method=solver), case.phi, km, kn, st, info=info)
    t1.record(); torch.cuda.synchronize()
    print(f"{cfg} {len(case.km)} pencils {solver}: invert {t0.elapsed_time(t1):.3f} ms, info max {int(info.max())}")
    y = torch.zeros_like(st)
    t0.record(); op.accumulate_batch(case.phi, km, kn, st, 0.0, y); t1.record(); torch.cuda.synchronize()
    print(f"accumulate {t0.elapsed_time(t1):.3f} ms")
import ctypes
_lib = ctypes.CDLL(os.path.join(ROOT, "suzerain_b200", "libsuzerain_b200.so"))
if hasattr(_lib, "szb_debug_pipe_prof"):
    buf = (ctypes.c_ulonglong * 16)()
    if _lib.szb_debug_pipe_prof(buf, 1) == 1:
        # three invert launches happened above per rep x 2 reps; report per pencil-panel
        n = 4 * len(case.km) * (5 * case.n // 5)
        names = ["P:lookahead", "P:load", "P:F", "P:publish", "P:endbar", "-", "-", "-",
                 "U:lookahead", "U:main", "U:barupd", "U:R", "U:A", "U:endbar"]
        print("phase clocks per panel:", {k: round(buf[i] / n) for i, k in enumerate(names) if k != "-"})
