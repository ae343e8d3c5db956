#!/usr/bin/env python
"""Where the host-pointer (e2e) substep spends its time: accumulate / invert separately, both solvers."""
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
hin = torch.empty((npen, 5, n), dtype=torch.complex128).pin_memory(); hin.copy_(a0)
hout = torch.zeros((5, npen, n), dtype=torch.complex128).pin_memory()
for solver in ("zgbsv", "zcgbsvx"):
    OH = sz.OperatorHybridIsothermal(op, wl.grid, sz.SolverSpec(method=solver))
    hin.copy_(a0)
    ta = ti = 0.0
    for i in range(5):
        pa, beta, pi = wl.phis(i)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        OH.accumulate_mass_plus_scaled_operator(pa, hin.numpy(), beta, hout.numpy(), npen * n)
        t1 = time.perf_counter()
        OH.invert_mass_plus_scaled_operator(pi, hin.numpy())
        t2 = time.perf_counter()
        if i >= 2:
            ta += t1 - t0; ti += t2 - t1
    print(f"{solver}: host accumulate {ta / 3 * 1e3:.2f} ms, host invert {ti / 3 * 1e3:.2f} ms")
    # device-resident invert of the same state, with iteration counts
    H = sz.OperatorHybridIsothermalDevice(op, wl.grid, sz.SolverSpec(method=solver), dev)
    a = a0.clone()
    iters = torch.zeros(H.nactive, dtype=torch.int32, device=dev)
    for rep in range(2):
        a.copy_(a0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        H.invert_mass_plus_scaled_operator(wl.phis(0)[2], a, iters=iters)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{solver}: device invert of the raw state {1e3 * (t1 - t0):.2f} ms, iters histogram {np.bincount(iters.cpu().numpy().clip(0))}")
