#!/usr/bin/env python
"""Wall-clock split of the whole-field host-pointer calls (accumulate / invert) used by bench.py's e2e."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import suzerain_b200 as sz
import bench
wl = bench.Workload("channel_192x96x192")
op = wl.make_imexop()
spec = sz.SolverSpec(method=(sys.argv[1] if len(sys.argv) > 1 else "zgbsv"))
OH = sz.OperatorHybridIsothermal(op, wl.grid, spec)
h = wl.host_state()
hin = torch.from_numpy(h).pin_memory(); hout = torch.zeros((5, wl.npencil, wl.Ny), dtype=torch.complex128).pin_memory()
a, b = hin.numpy(), hout.numpy(); fs = wl.npencil * wl.Ny
for it in range(4):
    pa, beta, pi = wl.phis(it)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    OH.accumulate_mass_plus_scaled_operator(pa, a, beta, b, fs)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    OH.invert_mass_plus_scaled_operator(pi, a)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"accumulate {1e3*(t1-t0):.2f} ms   invert {1e3*(t2-t1):.2f} ms")
