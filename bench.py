#!/usr/bin/env python
"""bench.py -- ns per grid point per SMR91 substep of the implicit wall-normal
operator path (BASELINE.json), on N B200s of one node.

One "step" = the linear-operator work of one SMR91 substep over every local
(kx,kz) pencil of the named grid (suzerain/lowstorage.hpp:1501-1515):

    tmp <- (M + dt*alpha_i L) a + chi*dt*zeta_i tmp      accumulate_mass_plus_scaled_operator
    a <-> tmp                                            b.exchange(a)  (pointer swap here)
    a   <- (M - dt*beta_i L)^-1 a                        invert_mass_plus_scaled_operator
                                                         (dealiased pencils zero-filled)

The nonlinear operator N (FFTs, transposes, pointwise physics) is outside the
hot path this repository rebuilds and is *not* in the timed region; the number
is the L-operator part of the substep.

Arms:
  default            device-resident state, kernels launched through the C ABI
                     (libsuzerain_b200.so); `value` is timed with CUDA events.
                     `e2e` repeats the same work through the HOST-pointer
                     whole-field entry points (H2D + D2H inside the timing).
  --impl reference   the reference's own C sources (oracle/_ref, OpenBLAS
                     LAPACK; MKL is not in the image) on the host cores.

Multi-GPU: wavenumber pencils are independent, so every rank owns its own
block of (kx,kz) pencils (weak scaling: the per-GPU block is the named grid's
whole wave space; the job's grid is N blocks side by side in z).  No
data-path collective; ranks meet only in the timing barrier.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="channel_192x96x192")
    ap.add_argument("--solver", default="zgbsv", choices=["zgbsv", "zcgbsvx"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------
class Workload:
    """Synthetic turbulent-channel state of one named grid."""

    def __init__(self, name, rank=0):
        import suzerain_b200 as sz
        from suzerain_b200 import synth
        self.name = name
        Nx, Ny, Nz, k, htdelta, one_sided = synth.CONFIGS[name]
        self.Nx, self.Ny, self.Nz, self.k, self.one_sided = Nx, Ny, Nz, k, one_sided
        self.Ly = 2.0
        bp = sz.htstretch_breakpoints(Ny, k, 0.0, self.Ly, htdelta)
        self.bop = sz.BsplineOp.from_breakpoints(k, bp)
        self.scenario = dict(synth.SCENARIO)
        self.refs = synth.reference_profiles(self.bop.greville(), self.Ly, self.scenario, one_sided)
        self.walls = synth.isothermal_walls(one_sided)
        self.nrbc = synth.nrbc_matrices() if one_sided else None
        self.grid = sz.wavegrid(Nx, Nz, synth.LX, synth.LZ)
        km, kn, act = sz.wavenumbers(self.grid)
        self.km, self.kn, self.act = km, kn, act
        self.npencil, self.nactive = len(km), int(act.sum())
        self.dt = synth.delta_t(self.Ly)
        self.chi = 1.0 / (self.grid.dNx * self.grid.dNz)
        self.gridpoints = Nx * Ny * Nz
        self.seed = synth.SEED + 1000 * rank
        self.synth = synth

    def host_state(self):
        """(npencil, 5, Ny) complex128, dealiased pencils holding garbage that
        invert must zero-fill."""
        return self.synth.state(self.km, self.kn, self.Ny, self.seed)

    def phis(self, i):
        s = self.synth
        i %= 3
        return (complex(self.dt * s.SMR91_ALPHA[i]), complex(self.chi * self.dt * s.SMR91_ZETA[i]),
                complex(-self.dt * s.SMR91_BETA[i]))

    def make_imexop(self):
        import suzerain_b200 as sz
        op = sz.ImexOp(self.bop)
        op.set_scenario(**self.scenario)
        op.set_refs(self.refs)
        op.set_isothermal(self.walls["enforce_lower"], self.walls["enforce_upper"],
                          self.walls["lower"], self.walls["upper"])
        if self.nrbc is not None:
            op.set_nrbc(*self.nrbc)
        return op

    def bc_dict(self):
        g, Ma = self.scenario["gamma"], self.scenario["Ma"]
        lo, up = self.walls["lower"], self.walls["upper"]
        ef = [w[0] / (g * (g - 1)) + Ma * Ma / 2 * (w[1] ** 2 + w[2] ** 2 + w[3] ** 2) for w in (lo, up)]
        return dict(enforce_lower=int(self.walls["enforce_lower"]),
                    enforce_upper=int(self.walls["enforce_upper"]),
                    E_factor=ef, vel_factor=[list(lo[1:]), list(up[1:])])


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# reference / CPU arm
# ---------------------------------------------------------------------------
def cpu_substep_time(wl: Workload, solver, cores, seconds, steps=1, warmup=0):
    """Times the reference's own per-pencil loop bodies (oracle/_ref: unmodified
    reference C + LAPACK from OpenBLAS) on a strided sample of the active
    pencils; returns (seconds per full-workload substep, sample description,
    kind)."""
    from oracle import ref as oref
    if not oref.available():
        raise RuntimeError("oracle/_ref/libsuzerain_ref.so is missing (run __graft_entry__.build())")
    P = oref.Problem(wl.bop, wl.scenario, wl.refs, wl.bc_dict(), wl.nrbc)
    km, kn = wl.km[wl.act], wl.kn[wl.act]
    # calibrate on a small sample, then size the timed sample to ~`seconds` per step
    ncal = min(len(km), 16 * cores)
    sel = np.linspace(0, len(km) - 1, ncal).astype(int)
    x = wl.synth.state(km[sel], kn[sel], wl.Ny, wl.seed).reshape(ncal, -1)
    pa, beta, pi = wl.phis(1)
    t0 = time.perf_counter()
    y = P.accumulate(pa, km[sel], kn[sel], x, beta=beta, y=x, nthreads=cores)
    P.invert(solver, pi, km[sel], kn[sel], y, nthreads=cores)
    per = (time.perf_counter() - t0) / ncal
    nsample = int(min(len(km), max(ncal, seconds / max(per, 1e-9))))
    sel = np.linspace(0, len(km) - 1, nsample).astype(int)
    x = wl.synth.state(km[sel], kn[sel], wl.Ny, wl.seed).reshape(nsample, -1)
    times = []
    for it in range(warmup + steps):
        pa, beta, pi = wl.phis(it)
        t0 = time.perf_counter()
        y = P.accumulate(pa, km[sel], kn[sel], x, beta=beta, y=x, nthreads=cores)
        r = P.invert(solver, pi, km[sel], kn[sel], y, nthreads=cores)
        dt = time.perf_counter() - t0
        assert r["info"] == 0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times)) * len(km) / nsample
    sample = (f"{nsample} of {len(km)} active pencils (evenly strided), accumulate+invert({solver}), "
              f"{cores} OpenMP threads one pencil each, OpenBLAS LAPACK (no MKL in image), scaled to the full grid")
    return t, sample, "reference"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.config)
    cores = os.cpu_count() or 1
    per_step = min(20.0, 150.0 / max(args.steps + args.warmup, 1))
    t, sample, kind = cpu_substep_time(wl, args.solver, cores, per_step, args.steps, args.warmup)
    # weak scaling: the job is N blocks; the host cores do them one after another
    ns = t * 1e9 / wl.gridpoints
    line = base_line(args, wl)
    line.update({"impl": "reference", "value": ns, "ms_per_step": t * 1e3 * args.gpus,
                 "cpu_baseline": {"value": ns, "unit": "ns/gridpoint/substep", "cores": cores,
                                  "kind": kind, "sample": sample},
                 "e2e": {"value": ns, "unit": "ns/gridpoint/substep", "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "dtype": "f64"})
    print_line(line)


def base_line(args, wl):
    return {"metric": "ns/gridpoint/SMR91 substep (implicit operator L: accumulate + invert)",
            "value": None, "unit": "ns/gridpoint/substep", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"perfect-gas channel {wl.Nx}x{wl.Ny}x{wl.Nz}, B-spline order {wl.k}, "
                                   f"per-GPU block = whole wave space ({wl.nactive} active of {wl.npencil} "
                                   f"stored (kx,kz) pencils, N={5 * wl.Ny})",
                       "name": wl.name, "solver": args.solver, "Ny": wl.Ny, "k": wl.k,
                       "active_pencils_per_gpu": wl.nactive, "stored_pencils_per_gpu": wl.npencil,
                       "gridpoints_per_gpu": wl.gridpoints,
                       "l2": "state (2 x %.0f MB per GPU) exceeds the 126 MB L2" % (wl.npencil * 5 * wl.Ny * 16 / 1e6),
                       "parallelism": f"(kx,kz) blocks x{args.gpus}, no data-path collective"}}


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """One process per GPU: run on (and first-touch pinned host memory from) the cores of the NUMA
    node the GPU hangs off, so that the host<->device copies of the e2e leg do not cross the
    socket interconnect.  Best effort; returns a description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return "numa node unknown"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += range(int(lo), int(hi or lo) + 1)
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return f"numa node {node}: no allowed cpus"
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception as e:                                   # noqa: BLE001
        return f"unbound ({type(e).__name__})"


def run_b200(args):
    import torch
    import torch.distributed as dist
    import suzerain_b200 as sz
    from suzerain_b200 import lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    binding = bind_to_gpu_numa_node(local) if world > 1 else "single process"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = L.load()
    wl = Workload(args.config, rank)
    op = wl.make_imexop()
    spec = sz.SolverSpec(method=args.solver)
    H = sz.OperatorHybridIsothermalDevice(op, wl.grid, spec, dev)
    h_state = wl.host_state()
    a = torch.from_numpy(h_state).to(dev)
    tmp = torch.zeros_like(a)
    stream = torch.cuda.current_stream()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    k_events = []          # (acc_start, acc_end/inv_start, inv_end) per timed step

    def substep(i, record=False):
        nonlocal a, tmp
        pa, beta, pi = wl.phis(i)
        if record:
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(stream)
        H.op.accumulate_batch(pa, H.km, H.kn, a, beta, tmp, index=H.active, stream=stream)
        if record:
            e1.record(stream)
        a, tmp = tmp, a
        H.invert_mass_plus_scaled_operator(pi, a, stream=stream)
        if record:
            e2.record(stream)
            k_events.append((e0, e1, e2))

    for i in range(args.warmup):
        substep(i)
    barrier()
    info = H.info.cpu().numpy()
    assert (info[:H.nactive] == 0).all(), "singular pencil in warm-up"
    launches0 = lib.szb_launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    t0, t1 = ev(), ev()
    t0.record(stream)
    for i in range(args.steps):
        substep(args.warmup + i, record=True)
    t1.record(stream)
    barrier()
    elapsed_ms = t0.elapsed_time(t1)
    launches = lib.szb_launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    assert torch.isfinite(torch.view_as_real(a)).all(), "state went non-finite"
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    acc_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1, _ in k_events]))
    inv_ms = float(np.mean([e1.elapsed_time(e2) for _, e1, e2 in k_events]))

    # ---- for information: the same invert under the reference's default --solver (zcgbsvx), on the evolved state ----
    default_solver_ms = None
    if args.solver == "zgbsv" and world == 1:
        try:
            Hd = sz.OperatorHybridIsothermalDevice(op, wl.grid, sz.SolverSpec(), dev)
            keep = a.clone()
            pi = wl.phis(0)[2]
            Hd.invert_mass_plus_scaled_operator(pi, a, stream=stream)          # warm-up / workspace
            a.copy_(keep)
            d0, d1 = ev(), ev()
            d0.record(stream)
            Hd.invert_mass_plus_scaled_operator(pi, a, stream=stream)
            d1.record(stream)
            torch.cuda.synchronize()
            default_solver_ms = d0.elapsed_time(d1)
            a.copy_(keep)
            del keep, Hd
        except Exception as e:                               # noqa: BLE001  (informational only)
            default_solver_ms = f"unavailable: {type(e).__name__}"

    # ---- e2e through the host-pointer whole-field entry points ----
    e2e = None
    if not args.no_e2e:
        OH = sz.OperatorHybridIsothermal(op, wl.grid, spec)
        n, npen = wl.Ny, wl.npencil
        hin = torch.from_numpy(h_state).pin_memory()
        hout = torch.zeros((5, npen, n), dtype=torch.complex128).pin_memory()
        hin_np, hout_np = hin.numpy(), hout.numpy()
        fs = npen * n

        def e2e_step(i):
            pa, beta, pi = wl.phis(i)
            OH.accumulate_mass_plus_scaled_operator(pa, hin_np, beta, hout_np, fs)
            OH.invert_mass_plus_scaled_operator(pi, hin_np)
        e2e_step(0)
        barrier()
        w0 = time.perf_counter()
        for i in range(args.e2e_steps):
            e2e_step(1 + i)
        barrier()
        e2e_s = (time.perf_counter() - w0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        # only active pencils cross the bus (accumulate: input + output up, output down;
        # invert: state up and down; dealiased pencils are zero-filled on the host)
        state_bytes = wl.nactive * 5 * n * 16
        e2e = {"value": e2e_s * 1e9 / (wl.gridpoints * world), "unit": "ns/gridpoint/substep",
               "h2d_bytes_per_step": 3 * state_bytes, "d2h_bytes_per_step": 2 * state_bytes,
               "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "api": "szb_operator_{accumulate,invert}_mass_plus_scaled_operator on pinned host state",
               "host_binding": binding}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    N = 5 * wl.Ny
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    inv_bytes = 2 * 16 * N * wl.nactive + 16 * N * (wl.npencil - wl.nactive)
    acc_bytes = (3 if True else 2) * 16 * N * wl.nactive
    inv_gbs = inv_bytes / (inv_ms * 1e-3) / 1e9
    acc_gbs = acc_bytes / (acc_ms * 1e-3) / 1e9
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl.name, {})
    except Exception:
        pass
    tr = lambda k: (traffic[k]["read"] + traffic[k]["write"]) if (k in traffic and args.solver == "zgbsv") else None
    KL = op.KL
    KU = op.KU
    lu_flop = 8.0 * N * KL * (KL + KU) + 8.0 * N * (2 * KL + KU)
    try:
        fp64_peak = float(json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")))["fp64_tflops"])
    except Exception:
        fp64_peak = 34.17
    line = base_line(args, wl)
    line.update({
        "value": ms_per_step * 1e6 / (wl.gridpoints * world), "ms_per_step": ms_per_step,
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"kernel": "invert (assemble + factor + solve, fused)", "bound": "hbm",
                     "achieved": inv_gbs, "peak": peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650",
                     "unit": "GB/s", "frac": inv_gbs / peak, "traffic": tr("invert_pipe"),
                     "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write per launch)",
                     "ms_per_launch": inv_ms, "algorithmic_bytes_per_launch": inv_bytes,
                     "note": "SURVEY 8d: the fused kernel only moves the 32 N bytes of state per system (assembly, "
                             "factors and U never touch HBM), so the HBM fraction is small by construction; the "
                             "binding roofline is the FP64 / latency one below (see DESIGN 3.1, 4.1)",
                     "fp64_gflops_upper": lu_flop * wl.nactive / (inv_ms * 1e-3) / 1e9,
                     # the kernel is an FP64 latency chain, not an HBM stream: the second roofline
                     "fp64": {"achieved": lu_flop * wl.nactive / (inv_ms * 1e-3) / 1e12, "peak": fp64_peak,
                              "unit": "TFLOP/s", "frac": lu_flop * wl.nactive / (inv_ms * 1e-3) / 1e12 / fp64_peak,
                              "peak_source": "profiles/r01_fp64_peak.json (tools/fp64_peak, DFMA loop on this pool)",
                              "flop_per_system": lu_flop}},
        "kernels": {"accumulate": {"ms": acc_ms, "GB/s": acc_gbs, "frac": acc_gbs / peak,
                                   "algorithmic_bytes": acc_bytes, "traffic": tr("accumulate")},
                    "invert": {"ms": inv_ms, "GB/s": inv_gbs, "frac": inv_gbs / peak},
                    "invert_default_solver_zcgbsvx": {"ms": default_solver_ms,
                                                      "note": "same state, reference's default --solver; not part of value"}},
    })
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        t, sample, kind = cpu_substep_time(wl, args.solver, cores, args.cpu_seconds)
        line["cpu_baseline"] = {"value": t * 1e9 / wl.gridpoints, "unit": "ns/gridpoint/substep",
                                "cores": cores, "kind": kind, "sample": sample}
    print_line(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # Libraries (NCCL's version banner, OpenMP notices) write to stdout; the contract is ONE
    # JSON line there.  Route fd 1 to stderr for the duration and print the line to the real one.
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    global print_line
    print_line = lambda obj: (out.write(json.dumps(obj) + "\n"), out.flush())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
