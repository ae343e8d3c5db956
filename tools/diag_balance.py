#!/usr/bin/env python
"""Per-pencil cost of the fused invert as a function of the wavenumber: shards of a big grid, a sample of each."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import suzerain_b200 as sz
from suzerain_b200 import lib as L
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "channel_1536x384x1152"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
lib = L.load()
buf = (ctypes.c_ulonglong * 30)()
for r in range(world):
    wl = bench.Workload(name, r, world, True)
    op = wl.make_imexop()
    km, kn = wl.km[wl.act], wl.kn[wl.act]
    sel = np.linspace(0, len(km) - 1, 3552).astype(int)
    kmt, knt = torch.from_numpy(km[sel]).to(dev), torch.from_numpy(kn[sel]).to(dev)
    x = torch.from_numpy(wl.synth.state(km[sel], kn[sel], wl.Ny, 1)).to(dev)
    info = torch.zeros(len(sel), dtype=torch.int32, device=dev)
    spec = sz.SolverSpec(method="zgbsv")
    pi = wl.phis(0)[2]
    op.invert_batch(spec, pi, kmt, knt, x.clone(), info=info); torch.cuda.synchronize()
    if hasattr(lib, "szb_debug_sync_prof"): lib.szb_debug_sync_prof(buf, 1)
    xx = x.clone(); torch.cuda.synchronize(); t0 = time.perf_counter()
    op.invert_batch(spec, pi, kmt, knt, xx, info=info); torch.cuda.synchronize(); t1 = time.perf_counter()
    frac = None
    if hasattr(lib, "szb_debug_sync_prof") and lib.szb_debug_sync_prof(buf, 1) == 1:
        frac = buf[6] / max(buf[7], 1)
    print(f"rank {r}: kz rows {wl.grid.dkbz}..{wl.grid.dkez} |kn| max {np.abs(kn).max():.1f}: {1e3 * (t1 - t0):.2f} ms for {len(sel)} pencils, exact-path panels {frac}")
