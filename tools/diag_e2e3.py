#!/usr/bin/env python
"""bench.py's e2e leg with the two host-pointer calls timed separately (same dataflow: invert sees the accumulate
output), per solver; iteration counts of the refinement on the same evolving state."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import suzerain_b200 as sz
import bench
wl = bench.Workload("channel_192x96x192")
op = wl.make_imexop()
dev = torch.device("cuda:0")
a0 = wl.device_state(dev)
npen, n = wl.npencil, wl.Ny
for solver in sys.argv[1:] or ("zgbsv", "zcgbsvx"):
    OH = sz.OperatorHybridIsothermal(op, wl.grid, sz.SolverSpec(method=solver))
    hin = torch.empty((npen, 5, n), dtype=torch.complex128).pin_memory(); hin.copy_(a0)
    hout = torch.zeros((5, npen, n), dtype=torch.complex128).pin_memory()
    fs = npen * n
    ta = ti = 0.0
    for i in range(5):
        pa, beta, pi = wl.phis(i)
        torch.cuda.synchronize(); w0 = time.perf_counter()
        OH.accumulate_mass_plus_scaled_operator(pa, hin.numpy(), beta, hout.numpy(), fs)
        w1 = time.perf_counter()
        tmp = hin.clone(); hin.copy_(hout.permute(1, 0, 2)); hout.copy_(tmp.permute(1, 0, 2))
        w2 = time.perf_counter()
        OH.invert_mass_plus_scaled_operator(pi, hin.numpy())
        w3 = time.perf_counter()
        if i >= 1:
            ta += w1 - w0; ti += w3 - w2
    print(f"{solver}: host accumulate {ta / 4 * 1e3:.2f} ms, host invert {ti / 4 * 1e3:.2f} ms  (SZB_HOST_STAGED={os.environ.get('SZB_HOST_STAGED', '1')})")
