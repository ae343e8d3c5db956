#!/usr/bin/env python
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
spec = sz.SolverSpec(method="zcgbsvx")
H = sz.OperatorHybridIsothermalDevice(op, wl.grid, spec, dev)
pi = wl.phis(0)[2]
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
a = a0.clone()
info = torch.zeros(H.nactive, dtype=torch.int32, device=dev)
def whole(stream=None):
    a.copy_(a0)
    H.op.invert_batch(spec, pi, H.km, H.kn, a, index=H.active, info=info, stream=stream)
print("whole, default stream: %.2f ms" % timed(whole))
s2 = torch.cuda.Stream()
print("whole, side stream   : %.2f ms" % timed(lambda: whole(s2)))
na = H.nactive
cuts = [0, na // 8, na - na // 8, na]
def chunks(stream=None):
    a.copy_(a0)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        H.op.invert_batch(spec, pi, H.km[lo:hi], H.kn[lo:hi], a, index=H.active[lo:hi], info=info[lo:hi], stream=stream)
print("3 chunks, default stream: %.2f ms" % timed(chunks))
print("3 chunks, side stream   : %.2f ms" % timed(lambda: chunks(s2)))
spec0 = sz.SolverSpec(method="zgbsv")
def chunks0(stream=None):
    a.copy_(a0)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        H.op.invert_batch(spec0, pi, H.km[lo:hi], H.kn[lo:hi], a, index=H.active[lo:hi], info=info[lo:hi], stream=stream)
print("zgbsv 3 chunks, side stream: %.2f ms" % timed(lambda: chunks0(s2)))
# repeated in-place inversion of the same state (what the e2e leg does with hin)
a.copy_(a0)
iters = torch.zeros(H.nactive, dtype=torch.int32, device=dev)
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    H.op.invert_batch(spec, wl.phis(rep)[2], H.km, H.kn, a, index=H.active, info=info, iters=iters)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("repeat %d: %.2f ms iters %s |a|max %.3g" % (rep, 1e3 * (t1 - t0), np.bincount(iters.cpu().numpy().clip(0)), a.abs().max().item()))
