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

# the other members of the bsplineop family: in place (one read, one write) and real pencils
ms = timed(lambda: sz.bsplineop_apply_batch(wl.bop, 0, 1.0, X2))
out["kernels"]["bop_apply_in_place(complex)"] = {"ms": ms, "GB/s": 2 * nbytes / ms / 1e6, "frac": 2 * nbytes / ms / 1e6 / peak}
XR, YR = torch.view_as_real(x).reshape(-1, Ny), torch.view_as_real(y).reshape(-1, Ny)
ms = timed(lambda: sz.bsplineop_accumulate_batch(wl.bop, 1, 1.0, XR, 0.5, YR))
out["kernels"]["bop_accumulate(real, beta!=0)"] = {"ms": ms, "GB/s": 3 * nbytes / ms / 1e6, "frac": 3 * nbytes / ms / 1e6 / peak}
# collect_references on the dealiased physical extent of the grid: five fields in, 42 x Ny out
pz, px = g.dNz, g.dNx
gen = torch.Generator(device=dev); gen.manual_seed(1)
sphys = torch.rand((5, Ny, pz, px), dtype=torch.float64, device=dev, generator=gen) + 1.0
sphys[0] += 5.0
scen = dict(Re=3000.0, Pr=0.7, Ma=1.5, alpha=0.0, gamma=1.4)
refs42 = torch.empty((Ny, 42), dtype=torch.float64, device=dev)
ms = timed(lambda: sz.collect_references(scen, 2.0 / 3.0, sphys, out=refs42))
pb = sphys.numel() * 8
out["kernels"]["collect_references"] = {"ms": ms, "GB/s": pb / ms / 1e6, "frac": pb / ms / 1e6 / peak,
                                        "points": int(Ny * pz * px), "note": "algorithmic bytes = the five physical fields"}

# invert under the other linearisation / solver specifications, all active pencils
act = np.flatnonzero(wl.act)
km = torch.from_numpy(wl.km[act]).to(dev); kn = torch.from_numpy(wl.kn[act]).to(dev)
st0 = torch.from_numpy(wl.host_state()[act]).to(dev)
info = torch.zeros(len(act), dtype=torch.int32, device=dev)
phi = wl.phis(0)[2]
out["invert"] = {"pencils": int(len(act)), "note": "ms per call over all active pencils, state resident"}
for lin, text, reps in (("rhome_xyz", "zgbsv", 5), ("rhome_xyz", "zcgbsvx", 3), ("rhome_xyz", "zgbsvx,equil=false", 2),
                        ("rhome_xyz", "zgbsvx,equil=true", 2), ("rhome_y", "zgbsv", 10), ("rhome_y", "zcgbsvx", 3)):
    op = wl.make_imexop().set_linearization(lin)
    spec = sz.SolverSpec.parse(text)
    st = st0.clone()

    def run():
        st.copy_(st0)
        op.invert_batch(spec, phi, km, kn, st, info=info)
    ms_copy = timed(lambda: st.copy_(st0), reps=10)
    ms = timed(run, reps=reps) - ms_copy
    assert int(info.max()) == 0
    out["invert"][f"{lin} {text}"] = {"ms": ms}
print(json.dumps(out))
