#!/usr/bin/env python
"""Throughput of the two wave-space building blocks of the nonlinear operator on the
bench grid: batched B-spline operator apply and diffwave, in GB/s of algorithmic traffic against
the measured HBM copy peak.  python tools/bench_aux.py [config]  -> one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import suzerain_b200 as sz
import bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "channel_192x96x192"
wl = bench.Workload(cfg)
dev = torch.device("cuda:0")
g = wl.grid
nz, nx, Ny = g.dkez - g.dkbz, g.dkex - g.dkbx, wl.Ny
nf = 5                                              # five scalar fields: 5 x 64 MB > L2
x = torch.randn((nf * nz, nx, Ny), dtype=torch.complex128, device=dev)
y = torch.randn_like(x)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


out = {"config": cfg, "elements": x.numel(), "peak_GB/s": peak, "kernels": {}}
nbytes = x.numel() * 16
X2, Y2 = x.view(-1, Ny), y.view(-1, Ny)
for name, beta, traffic in (("bop_apply(beta=0)", 0.0, 2), ("bop_accumulate(beta!=0)", 0.5, 3)):
    ms = timed(lambda: sz.bsplineop_accumulate_complex_batch(wl.bop, 1, 1.0, X2, beta, Y2))
    out["kernels"][name] = {"ms": ms, "GB/s": traffic * nbytes / ms / 1e6, "frac": traffic * nbytes / ms / 1e6 / peak}
# diffwave works on one field's wave space at a time
xs = [x[i * nz:(i + 1) * nz] for i in range(nf)]
ys = [y[i * nz:(i + 1) * nz] for i in range(nf)]
ms = timed(lambda: [sz.diffwave_accumulate(1, 0, 1.0, a, 0.5, b, g) for a, b in zip(xs, ys)])
out["kernels"]["diffwave_accumulate"] = {"ms": ms, "GB/s": 3 * nbytes / ms / 1e6, "frac": 3 * nbytes / ms / 1e6 / peak}
ms = timed(lambda: [sz.diffwave_apply(1, 1, 1.0, a, g) for a in xs])
# apply reads and writes kept pencils, only writes dealiased ones
out["kernels"]["diffwave_apply"] = {"ms": ms, "GB/s": 2 * nbytes / ms / 1e6, "frac": 2 * nbytes / ms / 1e6 / peak,
                                    "note": "upper bound on bytes: dealiased pencils are written, not read"}
print(json.dumps(out))
