#!/usr/bin/env python
"""BASELINE.json configs[4] / SURVEY 8d "batched-sweep config": the fused invert (assemble + factor +
solve, zgbsv) over batches of 4k .. 1M wavenumbers, Ny 96 .. 768, B-spline order 4 .. 10, operators
built by the real assembly (channel profiles, synthetic wavenumber lists), right hand sides random.
Per case: ms, ns per system, FP64 GFLOP/s (8 N KL (KL+KU) + 8 N (2KL+KU) flop per system, SURVEY 8a
row 9), the GB/s a pre-assembled API would have had to read (16 N LD bytes per system + 2 x 16 N),
and the reference's own C path (oracle/_ref, OpenBLAS LAPACK) on a bounded sample of the same batch
with all host threads.  python tools/sweep_invert.py [--quick] [--no-cpu]  -> one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import suzerain_b200 as sz
import parity_common as pc

quick = "--quick" in sys.argv
no_cpu = "--no-cpu" in sys.argv
dev = torch.device("cuda:0")
cases = [(8, 96, 4096), (8, 96, 65536), (8, 96, 1048576), (4, 96, 65536), (6, 96, 65536), (10, 96, 65536),
         (8, 192, 4096), (8, 192, 65536), (8, 384, 4096), (8, 384, 65536), (4, 384, 65536), (10, 384, 65536),
         (8, 768, 4096), (8, 768, 65536)]
if quick:
    cases = [(8, 96, 4096), (6, 192, 4096), (8, 768, 1024)]
threads = os.cpu_count() or 1
rows = []
for k, Ny, npen in cases:
    case = pc.make_case("channel_192x96x192", Ny=Ny, k=k, npencils=64)       # operators, profiles, phi
    rng = np.random.default_rng(7)
    km = np.concatenate([[0.0], rng.integers(-40, 41, npen - 1) * (2 * np.pi / pc.synth.LX)])
    kn = np.concatenate([[0.0], rng.integers(-40, 41, npen - 1) * (2 * np.pi / pc.synth.LZ)])
    op = pc.make_imexop(case)
    N, KL, KU, LD = op.N, op.KL, op.KU, op.LD
    dkm, dkn = torch.from_numpy(km).to(dev), torch.from_numpy(kn).to(dev)
    g = torch.Generator(device=dev); g.manual_seed(11)
    st0 = torch.view_as_complex(torch.randn((npen, 5, Ny, 2), dtype=torch.float64, device=dev, generator=g))
    st = st0.clone()
    info = torch.zeros(npen, dtype=torch.int32, device=dev)
    spec = sz.SolverSpec(method="zgbsv")
    op.invert_batch(spec, case.phi, dkm, dkn, st, info=info)                  # warm-up, workspace
    assert int(info.abs().max()) == 0, (k, Ny, npen)
    reps = 3 if npen * Ny <= 65536 * 384 else 1
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    for _ in range(reps):
        st.copy_(st0)
        t0.record(); op.invert_batch(spec, case.phi, dkm, dkn, st, info=info); t1.record()
        torch.cuda.synchronize()
        ms += t0.elapsed_time(t1) / reps
    flop = 8.0 * N * KL * (KL + KU) + 8.0 * N * (2 * KL + KU)
    row = {"k": k, "Ny": Ny, "N": N, "KL": KL, "systems": npen, "ms": ms, "ns_per_system": ms * 1e6 / npen,
           "fp64_GFLOP/s_upper": flop * npen / ms / 1e6,
           "preassembled_equiv_GB/s": (16.0 * N * LD + 32.0 * N) * npen / ms / 1e6,
           "state_GB/s": 32.0 * N * npen / ms / 1e6}
    if not no_cpu:
        ns = min(npen, 64 * threads)
        P = pc.oracle_problem(case, "ref")
        x = st0[:ns].cpu().numpy().reshape(ns, -1)
        P.invert("zgbsv", case.phi, km[:threads], kn[:threads], x[:threads], nthreads=threads)      # warm the thread pool
        w0 = time.perf_counter()
        r = P.invert("zgbsv", case.phi, km[:ns], kn[:ns], x, nthreads=threads)
        cpu_s = time.perf_counter() - w0
        assert r["info"] == 0
        got = st[:ns].cpu().numpy().reshape(ns, -1)
        row["cpu_ns_per_system"] = cpu_s * 1e9 / ns
        row["cpu_threads"] = threads
        row["cpu_sample"] = ns
        row["relmax_vs_reference_on_sample"] = float(np.abs(got - r["x"]).max() / np.abs(r["x"]).max())
        row["speedup_vs_host"] = row["cpu_ns_per_system"] / row["ns_per_system"]
    rows.append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
    del st, st0
    torch.cuda.empty_cache()
print(json.dumps({"sweep": rows, "solver": "zgbsv (fused assemble + factor + solve)", "device": torch.cuda.get_device_name(0)}))
