#!/usr/bin/env python
"""Host<->device copy bandwidth from pinned memory first-touched on each NUMA node's cores (is the e2e leg crossing sockets?)."""
import os, subprocess, sys, time
import torch
print(subprocess.run("nvidia-smi topo -m | head -14; lscpu | grep -i -E 'numa|socket|model name'; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c",
                     shell=True, capture_output=True, text=True).stdout)
print("affinity", len(os.sched_getaffinity(0)), "cpus")
nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
dev = torch.device("cuda", 0)
all_cpus = os.sched_getaffinity(0)
def cpulist(node):
    cpus = []
    for part in open(f"/sys/devices/system/node/{node}/cpulist").read().strip().split(","):
        lo, _, hi = part.partition("-"); cpus += range(int(lo), int(hi or lo) + 1)
    return [c for c in cpus if c in all_cpus]
nb = 256 << 20
d = torch.empty(nb, dtype=torch.uint8, device=dev); d2 = torch.empty(nb, dtype=torch.uint8, device=dev)
for node in nodes + ["all"]:
    cpus = cpulist(node) if node != "all" else list(all_cpus)
    if not cpus: continue
    os.sched_setaffinity(0, cpus)
    h = torch.empty(nb, dtype=torch.uint8).pin_memory(); h.fill_(1)
    h2 = torch.empty(nb, dtype=torch.uint8).pin_memory(); h2.fill_(2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    up = t(lambda: d.copy_(h, non_blocking=True)); down = t(lambda: h.copy_(d, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    bi = t(both)
    print(f"{node}: {len(cpus)} cpus  H2D {nb/up/1e9:.1f} GB/s  D2H {nb/down/1e9:.1f} GB/s  both at once {nb/bi/1e9:.1f} GB/s each way")
    del h, h2
