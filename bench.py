#!/usr/bin/env python
"""bench.py -- ns per grid point per SMR91 substep of the implicit wall-normal
operator path (BASELINE.json), on N B200s of one node.

One "step" = the linear-operator work of one SMR91 substep over every local
(kx,kz) pencil of the named grid (suzerain/lowstorage.hpp:1501-1515):

    b <- (M + dt*alpha_i L) a + chi*dt*zeta_i b          accumulate_mass_plus_scaled_operator
                                                         (a interleaved, b contiguous as in the reference)
    a <-> b                                              b.exchange(a)  (layout-converting swap kernel)
    a <- (M - dt*beta_i L)^-1 a                          invert_mass_plus_scaled_operator
                                                         (dealiased pencils zero-filled)
  + once per SMR91 step (every third substep): the reference profiles a sharded stepper sums
    over ranks (MPI_Allreduce of apps/perfect/perfect.cpp:1397 -> ncclAllReduce of the 42 x Ny
    block), the step-size candidates (MPI_Allreduce MIN of suzerain/support/driver_base.cpp:2052)
    and the refresh of the operator's profile table from the reduced device buffer.

The nonlinear operator N (FFTs, transposes, pointwise physics) is outside the
hot path this repository rebuilds and is *not* in the timed region; the number
is the L-operator part of the substep.

Grids (BASELINE.json configs): 1 GPU: channel_192x96x192; 2 and 4 GPUs: bl_1024x256x512;
8 GPUs: channel_1536x384x1152 -- each ONE grid sharded over the ranks by contiguous kz blocks
balanced on active pencils (suzerain_b200/shard.py, the reference's own decomposition with
whole wall-normal pencils per rank, suzerain/pencil_grid.cpp:118-157).  `--scaling weak` runs
the old layout instead (every rank owns a whole copy of the named grid).

Solver: the reference's default --solver (zcgbsvx,reuse=false,aiter=1,siter=-1,diter=5,tolsc=0,
suzerain/specification_zgbsv.cpp:46-54) is the headline; the same steps are also timed under
zgbsv (`solvers` in the JSON line), each with its own CPU number.

Arms:
  default            device-resident state, kernels launched through the C ABI
                     (libsuzerain_b200.so); `value` is timed with CUDA events.
                     `e2e` repeats the same work through the HOST-pointer
                     whole-field entry points (H2D + D2H inside the timing).
  --impl reference   the reference's own C sources (oracle/_ref, OpenBLAS
                     LAPACK; MKL is not in the image) on the host cores, rank 0 only.
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


DEFAULT_SOLVER = "zcgbsvx"          # the reference's default --solver (specification_zgbsv.cpp:46-54)
GRID_FOR_GPUS = {1: "channel_192x96x192", 2: "bl_1024x256x512", 4: "bl_1024x256x512",
                 8: "channel_1536x384x1152"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, help="named grid; default: the BASELINE.json grid for --gpus")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: ONE named grid sharded over the ranks; weak: a whole copy per rank")
    ap.add_argument("--balance", default="cost", choices=["cost", "active"],
                    help="sharding weights: measured row cost (default) or active pencils per row")
    ap.add_argument("--solver", default=DEFAULT_SOLVER, choices=["zgbsv", "zcgbsvx"])
    ap.add_argument("--no-second-solver", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.config is None:
        args.config = GRID_FOR_GPUS.get(args.gpus, "channel_1536x384x1152" if args.gpus > 4 else "bl_1024x256x512")
    if args.gpus == 1:
        args.scaling = "strong"          # one rank owns the whole grid either way
    return args


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------
class Workload:
    """Synthetic turbulent-channel / boundary-layer operator and state of one named grid,
    or of this rank's shard of it."""

    def __init__(self, name, rank=0, world=1, sharded=True, row_weights=None):
        import suzerain_b200 as sz
        from suzerain_b200 import synth, shard
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
        self.full_grid = sz.wavegrid(Nx, Nz, synth.LX, synth.LZ)
        self.sharded = sharded and world > 1
        self.grid = shard.shard_wavegrid(self.full_grid, rank, world, row_weights) if self.sharded else self.full_grid
        km, kn, act = sz.wavenumbers(self.grid)
        self.km, self.kn, self.act = km, kn, act
        self.npencil, self.nactive = len(km), int(act.sum())
        if self.sharded:
            self.total_active = int(shard.active_rows(self.full_grid).sum())
            self.total_pencils = (self.full_grid.dkex - self.full_grid.dkbx) * (self.full_grid.dkez - self.full_grid.dkbz)
            self.job_gridpoints = Nx * Ny * Nz
        else:
            self.total_active, self.total_pencils = self.nactive * world, self.npencil * world
            self.job_gridpoints = Nx * Ny * Nz * world
        self.dt = synth.delta_t(self.Ly)
        self.chi = 1.0 / (self.full_grid.dNx * self.full_grid.dNz)
        self.gridpoints = Nx * Ny * Nz
        self.seed = synth.SEED + 1000 * rank
        self.synth = synth

    def host_state(self):
        """(npencil, 5, Ny) complex128, dealiased pencils holding garbage that
        invert must zero-fill."""
        return self.synth.state(self.km, self.kn, self.Ny, self.seed)

    def device_state(self, dev):
        """The same law generated on the device (the large grids: no multi-GB host arrays)."""
        import torch
        g = torch.Generator(device=dev)
        g.manual_seed(self.seed)
        x = torch.empty((self.npencil, 5, self.Ny), dtype=torch.complex128, device=dev)
        xr = torch.view_as_real(x)
        chunk = max(1, (64 << 20) // (5 * self.Ny * 16))
        km = torch.from_numpy(self.km).to(dev); kn = torch.from_numpy(self.kn).to(dev)
        for p0 in range(0, self.npencil, chunk):
            p1 = min(self.npencil, p0 + chunk)
            amp = (1.0 + km[p0:p1] ** 2 + kn[p0:p1] ** 2) ** (-5.0 / 6.0)
            xr[p0:p1] = torch.randn((p1 - p0, 5, self.Ny, 2), dtype=torch.float64, device=dev, generator=g) \
                * amp[:, None, None, None]
        zero = (km == 0) & (kn == 0)
        if bool(zero.any()):
            xr[zero, :, :, 1] = 0.0                                  # the mean mode is real
        return x

    def references_block(self, dev):
        """The reference's 42 x Ny `references` block (apps/perfect/references.hpp:82-125), column-major:
        (Ny, 42) C-contiguous; rows q::u .. q::e_deltarho hold the 26 profiles, the others are zero here."""
        import torch
        blk = np.zeros((self.Ny, 42))
        blk[:, 5:31] = self.refs.T
        return torch.from_numpy(np.ascontiguousarray(blk)).to(dev)

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
    pencils of the WHOLE named grid; returns (seconds per full-grid substep -- the sample's
    time scaled by active pencils --, measured seconds per sample step, sample description, kind)."""
    from oracle import ref as oref
    if not oref.available():
        raise RuntimeError("oracle/_ref/libsuzerain_ref.so is missing (run __graft_entry__.build())")
    P = oref.Problem(wl.bop, wl.scenario, wl.refs, wl.bc_dict(), wl.nrbc)
    km, kn = wl.km[wl.act], wl.kn[wl.act]
    # calibrate on a small sample, then size the timed sample to ~`seconds` per step
    ncal = min(len(km), 8 * cores)
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
    t_sample = float(np.mean(times))
    t = t_sample * len(km) / nsample
    sample = (f"{nsample} of {len(km)} active pencils of {wl.name} (evenly strided), accumulate+invert({solver}), "
              f"{cores} OpenMP threads one pencil each, OpenBLAS LAPACK (no MKL in image), scaled to the full grid")
    return t, t_sample, sample, "reference"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.config)                        # the whole named grid
    cores = os.cpu_count() or 1
    per_step = min(15.0, 120.0 / max(args.steps + args.warmup, 1))
    t, t_sample, sample, kind = cpu_substep_time(wl, args.solver, cores, per_step, args.steps, args.warmup)
    # strong: the job is the named grid once; weak: N copies one after another on the same host cores --
    # the same ns per grid point either way
    ns = t * 1e9 / wl.gridpoints
    line = base_line(args, wl, args.gpus)
    line.update({"impl": "reference", "value": ns,
                 # measured: the bounded sample each timed step actually ran
                 "ms_per_step": t_sample * 1e3,
                 "ms_per_step_note": "measured duration of one step over the bounded sample; `value` scales it to the job",
                 "ms_per_step_whole_job_extrapolated": t * 1e3 * (args.gpus if args.scaling == "weak" else 1),
                 "cpu_baseline": {"value": ns, "unit": "ns/gridpoint/substep", "cores": cores,
                                  "kind": kind, "sample": sample},
                 "e2e": {"value": ns, "unit": "ns/gridpoint/substep", "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "dtype": "f64"})
    print_line(line)


SOLVER_SPEC_TEXT = {"zgbsv": "zgbsv",
                    "zcgbsvx": "zcgbsvx,reuse=false,aiter=1,siter=-1,diter=5,tolsc=0 (the reference's default --solver)"}


def base_line(args, wl, world):
    sharded = args.scaling == "strong" and world > 1
    kind = "boundary layer (one-sided, NRBC)" if wl.one_sided else "perfect-gas channel"
    if world == 1:
        part = "one rank owns the whole wave space"
    elif sharded:
        part = (f"ONE grid sharded over {world} ranks by contiguous kz blocks balanced on "
                f"{'measured row cost' if args.balance == 'cost' else 'active pencils'} "
                f"(whole wall-normal pencils per rank), no collective inside L; per SMR91 step one ncclAllReduce(SUM) "
                f"of the 42 x Ny reference profiles and one ncclAllReduce(MIN) of 12 step-size candidates")
    else:
        part = f"{world} replicas: every rank owns a whole copy of the grid, no data-path collective"
    return {"metric": "ns/gridpoint/SMR91 substep (implicit operator L: accumulate + exchange + invert)",
            "value": None, "unit": "ns/gridpoint/substep", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": False,
            "scaling": "strong" if (sharded or world == 1) else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{kind} {wl.Nx}x{wl.Ny}x{wl.Nz}, B-spline order {wl.k}: "
                                   f"{wl.total_active if hasattr(wl, 'total_active') else wl.nactive} active of "
                                   f"{wl.total_pencils if hasattr(wl, 'total_pencils') else wl.npencil} stored (kx,kz) "
                                   f"pencils, N={5 * wl.Ny} per pencil; {part}",
                       "name": wl.name, "solver": SOLVER_SPEC_TEXT[args.solver], "Ny": wl.Ny, "k": wl.k,
                       "active_pencils_this_rank": wl.nactive, "stored_pencils_this_rank": wl.npencil,
                       "job_gridpoints": wl.job_gridpoints,
                       "l2": "state (2 x %.0f MB on this rank) exceeds the 126 MB L2" % (wl.npencil * 5 * wl.Ny * 16 / 1e6),
                       "parallelism": part}}


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


def row_cost_weights(wl_full, op, dev, nbuckets=24, per_bucket=1184):
    """Measured cost of every kz row of the whole grid for the sharding (suzerain_b200/shard.py): the fused
    invert is timed on a sample of the active pencils of `nbuckets` groups of rows; a row weighs its active
    pencils times its group's time per pencil.  Run on rank 0 and broadcast, so that all ranks cut alike."""
    import torch
    import suzerain_b200 as sz
    from suzerain_b200 import shard
    g = wl_full.full_grid
    nx = g.dkex - g.dkbx
    nrows = g.dkez - g.dkbz
    act_rows = shard.active_rows(g).astype(np.float64)
    km = wl_full.km.reshape(nrows, nx); kn = wl_full.kn.reshape(nrows, nx); act = wl_full.act.reshape(nrows, nx)
    spec = sz.SolverSpec(method="zgbsv")
    pi = wl_full.phis(0)[2]
    cost = np.ones(nrows)
    edges = np.linspace(0, nrows, nbuckets + 1).astype(int)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for b in range(nbuckets):
        r0, r1 = edges[b], edges[b + 1]
        m = act[r0:r1]
        if not m.any():
            continue
        kmb, knb = km[r0:r1][m], kn[r0:r1][m]
        sel = np.linspace(0, len(kmb) - 1, min(per_bucket, len(kmb))).astype(int)
        kmt, knt = torch.from_numpy(kmb[sel]).to(dev), torch.from_numpy(knb[sel]).to(dev)
        x = torch.from_numpy(wl_full.synth.state(kmb[sel], knb[sel], wl_full.Ny, 7)).to(dev)
        info = torch.zeros(len(sel), dtype=torch.int32, device=dev)
        best = None
        for rep in range(3):
            xx = x.clone()
            e0, e1 = ev(), ev()
            e0.record(); op.invert_batch(spec, pi, kmt, knt, xx, info=info); e1.record()
            torch.cuda.synchronize()
            if rep:
                best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
        cost[r0:r1] = best / len(sel)
    cost /= cost[cost > 0].min() if (cost > 0).any() else 1.0
    return act_rows * cost


def mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return None


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
    sharded = args.scaling == "strong"
    weights = None
    balance = "active pencils"
    if sharded and world > 1 and args.balance == "cost":
        # static load balancing on measured row costs: rank 0 probes, everybody cuts alike
        wl_full = Workload(args.config)
        op0 = wl_full.make_imexop()
        w = torch.zeros(wl_full.full_grid.dkez - wl_full.full_grid.dkbz, dtype=torch.float64, device=dev)
        if rank == 0:
            w.copy_(torch.from_numpy(row_cost_weights(wl_full, op0, dev)))
        dist.broadcast(w, src=0)
        weights = w.cpu().numpy()
        balance = "measured row cost (fused invert timed on a sample of every group of kz rows by rank 0, broadcast)"
        del op0, wl_full
    wl = Workload(args.config, rank, world, sharded, weights)
    op = wl.make_imexop()
    n, npen = wl.Ny, wl.npencil
    stream = torch.cuda.current_stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)

    # device state: a interleaved (npencil, 5, Ny), b contiguous (5, npencil, Ny) as in the reference
    # (suzerain/storage.hpp:235-236,267-268)
    a0 = wl.device_state(dev)
    refs_blk = wl.references_block(dev)              # this rank's contribution: 1/world of the profiles
    refs_part = refs_blk / world
    dt_cand = torch.full((12,), wl.dt, dtype=torch.float64, device=dev)

    def step_collectives(H):
        """Once per SMR91 step: what a sharded stepper exchanges outside L."""
        r = refs_part.clone()
        d = dt_cand.clone()
        if world > 1:
            dist.all_reduce(r, op=dist.ReduceOp.SUM)           # apps/perfect/perfect.cpp:1397
            dist.all_reduce(d, op=dist.ReduceOp.MIN)           # suzerain/support/driver_base.cpp:2052
        H.op.set_refs_device(r, stream=stream)
        return r

    def timed_run(solver, steps, warmup):
        """`warmup` untimed + `steps` timed substeps under `solver`; returns per-step ms (max over ranks),
        mean kernel times and the launch count of the timed region."""
        spec = sz.SolverSpec(method=solver)
        H = sz.OperatorHybridIsothermalDevice(op, wl.grid, spec, dev)
        a = a0.clone()
        b = torch.zeros((5, npen, n), dtype=torch.complex128, device=dev)
        k_events = []

        def substep(i, record=False):
            pa, beta, pi = wl.phis(i)
            if i % 3 == 0:
                step_collectives(H)
            if record:
                e0, e1, e2, e3 = ev(), ev(), ev(), ev()
                e0.record(stream)
            H.accumulate_mass_plus_scaled_operator(pa, a, beta, b, stream=stream)
            if record:
                e1.record(stream)
            H.exchange(a, b, stream=stream)
            if record:
                e2.record(stream)
            H.invert_mass_plus_scaled_operator(pi, a, stream=stream)
            if record:
                e3.record(stream)
                k_events.append((e0, e1, e2, e3))

        for i in range(warmup):
            substep(i)
        barrier()
        info = H.info.cpu().numpy()
        assert (info[:H.nactive] == 0).all(), "singular pencil in warm-up"
        launches0 = lib.szb_launch_count()
        barrier()
        t0, t1 = ev(), ev()
        t0.record(stream)
        for i in range(steps):
            substep(warmup + i, record=True)
        t1.record(stream)
        barrier()
        ms = t0.elapsed_time(t1)
        launches = lib.szb_launch_count() - launches0
        assert torch.isfinite(torch.view_as_real(a)).all(), "state went non-finite"
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        acc = float(np.mean([e0.elapsed_time(e1) for e0, e1, _, _ in k_events]))
        exc = float(np.mean([e1.elapsed_time(e2) for _, e1, e2, _ in k_events]))
        inv = float(np.mean([e2.elapsed_time(e3) for _, _, e2, e3 in k_events]))
        del a, b, H
        return ms / steps, acc, exc, inv, int(launches)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_per_step, acc_ms, exc_ms, inv_ms, launches = timed_run(args.solver, args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None

    other = "zgbsv" if args.solver == "zcgbsvx" else "zcgbsvx"
    second = None
    if not args.no_second_solver:
        second = timed_run(other, args.steps, args.warmup)
    # the fused invert kernel alone (zgbsv = one launch of it) for the roofline entry
    if args.solver == "zgbsv":
        fused_ms = inv_ms
    elif second is not None:
        fused_ms = second[3]
    else:
        fused_ms = None

    # ---- e2e through the host-pointer whole-field entry points ----
    e2e = None
    if not args.no_e2e:
        spec = sz.SolverSpec(method=args.solver)
        OH = sz.OperatorHybridIsothermal(op, wl.grid, spec)
        state_bytes_all = npen * 5 * n * 16
        avail = mem_available_bytes()
        note = None
        if avail is not None and 2 * state_bytes_all * world > 0.5 * avail:
            note = f"skipped: pinned host state 2 x {state_bytes_all / 1e9:.1f} GB x {world} ranks exceeds half of MemAvailable"
        if note is None:
            hin = torch.empty((npen, 5, n), dtype=torch.complex128).pin_memory()
            hin.copy_(a0)
            hout = torch.zeros((5, npen, n), dtype=torch.complex128).pin_memory()
            hin_np, hout_np = hin.numpy(), hout.numpy()
            fs = npen * n

            def e2e_step(i):
                """The substep's two operator calls on host state, timed; the caller's own host-side
                b.exchange(a) between them (reference code, in neither arm's timing) is done untimed so that
                invert sees the accumulate output as in lowstorage::step -- inverting the same state over
                and over would blow it up and with it the refinement count of zcgbsvx."""
                pa, beta, pi = wl.phis(i)
                torch.cuda.synchronize()
                w0 = time.perf_counter()
                OH.accumulate_mass_plus_scaled_operator(pa, hin_np, beta, hout_np, fs)
                w1 = time.perf_counter()
                tmp = hin.clone()
                hin.copy_(hout.permute(1, 0, 2))                 # a <-> b, layouts converted
                hout.copy_(tmp.permute(1, 0, 2))
                w2 = time.perf_counter()
                OH.invert_mass_plus_scaled_operator(pi, hin_np)
                return (w1 - w0) + (time.perf_counter() - w2)
            e2e_step(0)
            barrier()
            e2e_s = 0.0
            for i in range(args.e2e_steps):
                e2e_s += e2e_step(1 + i)
                barrier()
            e2e_s /= args.e2e_steps
            if world > 1:
                t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t.item())
            # only active pencils cross the bus (accumulate: input + output up, output down;
            # invert: state up and down; dealiased pencils are zero-filled on the host)
            state_bytes = wl.nactive * 5 * n * 16
            e2e = {"value": e2e_s * 1e9 / wl.job_gridpoints, "unit": "ns/gridpoint/substep",
                   "h2d_bytes_per_step": 3 * state_bytes, "d2h_bytes_per_step": 2 * state_bytes,
                   "bytes_note": "per rank", "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
                   "api": "szb_operator_{accumulate,invert}_mass_plus_scaled_operator on pinned host state (wall time of "
                          "the two calls; the caller's host-side exchange between them is outside, as in the reference arm), "
                          + SOLVER_SPEC_TEXT[args.solver],
                   "host_binding": binding}
        else:
            e2e = {"value": None, "unit": "ns/gridpoint/substep", "note": note}

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
    acc_bytes = 3 * 16 * N * wl.nactive
    exc_bytes = 4 * 16 * N * wl.npencil
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl.name, {})
    except Exception:
        pass
    tr = lambda k: (traffic[k]["read"] + traffic[k]["write"]) if k in traffic else None
    KL, KU = op.KL, op.KU
    lu_flop = 8.0 * N * KL * (KL + KU) + 8.0 * N * (2 * KL + KU)
    try:
        fp64_peak = float(json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")))["fp64_tflops"])
    except Exception:
        fp64_peak = 34.17
    line = base_line(args, wl, world)
    if world > 1 and sharded:
        line["config"]["balance"] = balance
    gbs = lambda nbytes, ms: nbytes / (ms * 1e-3) / 1e9
    kernels = {"accumulate": {"ms": acc_ms, "GB/s": gbs(acc_bytes, acc_ms), "frac": gbs(acc_bytes, acc_ms) / peak,
                              "algorithmic_bytes": acc_bytes, "traffic": tr("accumulate")},
               "exchange": {"ms": exc_ms, "GB/s": gbs(exc_bytes, exc_ms), "frac": gbs(exc_bytes, exc_ms) / peak,
                            "algorithmic_bytes": exc_bytes},
               "invert": {"ms": inv_ms, "solver": args.solver, "GB/s": gbs(inv_bytes, inv_ms),
                          "frac": gbs(inv_bytes, inv_ms) / peak}}
    roof = None
    if fused_ms is not None:
        tf = lu_flop * wl.nactive / (fused_ms * 1e-3) / 1e12
        roof = {"kernel": "invert_sync_kernel (assemble + factor + solve, fused; the zgbsv invert is one launch of it)",
                "bound": "hbm", "achieved": gbs(inv_bytes, fused_ms), "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650",
                "unit": "GB/s", "frac": gbs(inv_bytes, fused_ms) / peak, "traffic": tr("invert_sync"),
                "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write per launch)",
                "ms_per_launch": fused_ms, "algorithmic_bytes_per_launch": inv_bytes,
                "note": "SURVEY 8d: the fused kernel only moves the 32 N bytes of state per system (assembly, "
                        "factors and U never touch HBM), so the HBM fraction is small by construction; the "
                        "binding rooflines are the FP64 one below and the shared-memory pipe (DESIGN 3.1)",
                "fp64": {"achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                         "peak_source": "profiles/r01_fp64_peak.json (tools/fp64_peak, DFMA loop on this pool)",
                         "flop_per_system": lu_flop}}
    line.update({
        "value": ms_per_step * 1e6 / wl.job_gridpoints, "ms_per_step": ms_per_step,
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "kernels": kernels,
    })
    solvers = {args.solver: {"value": ms_per_step * 1e6 / wl.job_gridpoints, "ms_per_step": ms_per_step,
                             "invert_ms": inv_ms, "spec": SOLVER_SPEC_TEXT[args.solver]}}
    if second is not None:
        solvers[other] = {"value": second[0] * 1e6 / wl.job_gridpoints, "ms_per_step": second[0],
                          "invert_ms": second[3], "spec": SOLVER_SPEC_TEXT[other]}
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        wl_full = wl
        for sv in solvers:
            t, t_sample, sample, kind = cpu_substep_time(wl_full, sv, cores, args.cpu_seconds / len(solvers))
            solvers[sv]["cpu"] = {"value": t * 1e9 / wl.gridpoints, "unit": "ns/gridpoint/substep", "cores": cores,
                                  "kind": kind, "sample": sample}
        line["cpu_baseline"] = solvers[args.solver]["cpu"]
    line["solvers"] = solvers
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
