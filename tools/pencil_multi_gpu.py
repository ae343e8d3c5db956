#!/usr/bin/env python
"""Multi-GPU check + timing of the pencil_grid replacement: run under torchrun (one rank per GPU).
Every rank builds the same global field, transforms its slab through pack -> NCCL all-to-all ->
finish, and compares with the numpy oracle's whole-field transform; then times both directions on
the bench grid (dealiased 288 x 96 x 288).  Prints one JSON line on rank 0."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import suzerain_b200 as sz
from oracle import pencil as op          # checker only

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
exchange = os.environ.get("SZB_PENCIL_EXCHANGE", "auto")
out = {"n_gpus": world, "exchange": exchange if world > 1 else "local"}
# ---- parity on a small ragged grid ----
dNx, Ny, dNz = 30, 11, 20
pg = sz.PencilGrid(dNx, Ny, dNz, exchange=exchange)
rng = np.random.default_rng(3)
phys = rng.standard_normal((Ny, dNz, dNx)); wave = op.physical_to_wave(phys)
ys, ye = pg.local_physical_start[1], pg.local_physical_end[1]
zs, ze = pg.local_wave_start[2], pg.local_wave_end[2]
buf = torch.zeros(pg.local_physical_storage(), dtype=torch.float64, device=dev)
pg.physical_view(buf).copy_(torch.from_numpy(phys[ys:ye]))
pg.transform_physical_to_wave(buf); torch.cuda.synchronize()
e1 = float(np.abs(pg.wave_view(buf).cpu().numpy() - wave[zs:ze]).max() / np.abs(wave).max())
pg.transform_wave_to_physical(buf); torch.cuda.synchronize()
e2 = float(np.abs(pg.physical_view(buf).cpu().numpy() - phys[ys:ye] * dNx * dNz).max() / (dNx * dNz * np.abs(phys).max()))
t = torch.tensor([e1, e2], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
out["parity_relmax"] = {"physical_to_wave": float(t[0]), "wave_to_physical_round_trip": float(t[1])}
assert float(t.max()) <= 1e-12, out
# ---- timing on the bench grid, five fields per call as the nonlinear operator transforms them ----
dNx, Ny, dNz = 288, 96, 288
pg = sz.PencilGrid(dNx, Ny, dNz, exchange=exchange)
bufs = [torch.randn(pg.local_physical_storage(), dtype=torch.float64, device=dev) for _ in range(5)]
def timed(fn, reps=10):
    for _ in range(3): fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / reps], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)
w2p = timed(lambda: [pg.transform_wave_to_physical(b) for b in bufs])
p2w = timed(lambda: [pg.transform_physical_to_wave(b) for b in bufs])
field_bytes = 8 * dNx * Ny * dNz                       # one real field, global
out["grid"] = [dNx, Ny, dNz]
out["exchange"] = pg.exchange
out["ms_five_fields"] = {"wave_to_physical": w2p, "physical_to_wave": p2w}
w2pm = timed(lambda: pg.transform_wave_to_physical_many(bufs))
p2wm = timed(lambda: pg.transform_physical_to_wave_many(bufs))
out["ms_five_fields_one_exchange"] = {"wave_to_physical": w2pm, "physical_to_wave": p2wm}
# algorithmic traffic of one transform: read the wave field, write the physical field (or back)
nxw = dNx // 2 + 1
alg = 5 * (16 * nxw * Ny * dNz + field_bytes)
out["algorithmic_GB/s_aggregate"] = {"wave_to_physical": alg / w2p / 1e6, "physical_to_wave": alg / p2w / 1e6}
ref = [b.clone() for b in bufs]
for b in bufs:
    pg.transform_physical_to_wave(b)          # make the buffers valid wave data of real fields
wave0 = [b.clone() for b in bufs]
pg.transform_wave_to_physical_many(bufs)
pg.transform_physical_to_wave_many(bufs)
torch.cuda.synchronize()
nx, ny, nz = pg.local_wave_extent
err = max(float((b[:2 * nx * ny * nz] / (dNx * dNz) - w[:2 * nx * ny * nz]).abs().max() / w[:2 * nx * ny * nz].abs().max())
          for b, w in zip(bufs, wave0)) if nx * ny * nz else 0.0
t = torch.tensor([err], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
out["many_round_trip_relmax"] = float(t)
assert float(t) <= 1e-12, out
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
