#!/usr/bin/env python
"""Phase clocks of the v5 invert kernel (invert_sync.cu) from a PROF=1 build:
    tools/build_variant.sh prof PROF=1
    SZB_LIB=suzerain_b200/variants/libprof.so python tools/prof_sync.py [config] [npencils]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import parity_common as pc
import suzerain_b200 as sz
from suzerain_b200 import lib as L

cfg = sys.argv[1] if len(sys.argv) > 1 else "channel_192x96x192"
npen = int(sys.argv[2]) if len(sys.argv) > 2 else 18336
case = pc.make_case(cfg, max_pencils=npen)
dev = torch.device("cuda:0")
op = pc.make_imexop(case)
km = torch.from_numpy(case.km).to(dev); kn = torch.from_numpy(case.kn).to(dev)
x0 = torch.from_numpy(case.x.copy()).to(dev)
info = torch.zeros(len(case.km), dtype=torch.int32, device=dev)
lib = L.load()
buf = (ctypes.c_ulonglong * 24)()
spec = sz.SolverSpec(method="zgbsv")
st = x0.clone(); op.invert_batch(spec, case.phi, km, kn, st, info=info); torch.cuda.synchronize()
has = hasattr(lib, "szb_debug_sync_prof") and lib.szb_debug_sync_prof(buf, 1) == 1
st = x0.clone()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); op.invert_batch(spec, case.phi, km, kn, st, info=info); t1.record(); torch.cuda.synchronize()
print(f"{cfg} {len(case.km)} pencils: invert {t0.elapsed_time(t1):.3f} ms, info max {int(info.max())}")
if has and lib.szb_debug_sync_prof(buf, 1) == 1:
    v = list(buf)
    npanel = max(v[7], 1)
    print("panels", v[7], "exact-path panels", v[6], f"({100.0 * v[6] / npanel:.1f} %)")
    names = ["P1 work", "B1 wait", "P2 work", "B2 wait", "P3 work", "B3 wait"]
    for base, who in ((0, "warp 0 (F1 | F2 | U)"), (8, "warp 2 (A  | X  | U)"), (16, "warp 5 (A  | coef | U)")):
        print(f"{who:24s}:", {k: round(v[base + i] / npanel) for i, k in enumerate(names)}, "sum", round(sum(v[base:base + 6]) / npanel))
