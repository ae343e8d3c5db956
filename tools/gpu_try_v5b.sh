#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2d}
timeout -s KILL 180 python - > gpurun_out/${TAG}_tiny.log 2>&1 <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, 'tests')
import parity_common as pc
dev = torch.device('cuda:0')
for name, kw in [("tiny_16x24x16", dict(max_pencils=24)), ("tiny_16x24x16", dict(max_pencils=20, k=8, Ny=40)),
                 ("channel_192x96x192", dict(max_pencils=40)), ("tiny_16x24x16", dict(max_pencils=30, phi=-5+3j, k=8, Ny=48)),
                 ("tiny_16x24x16", dict(max_pencils=20, k=10, Ny=48)), ("tiny_16x24x16", dict(max_pencils=20, k=4, Ny=40))]:
    case = pc.make_case(name, **kw)
    got = pc.gpu_invert(case, "zgbsv", dev)
    want = pc.oracle_invert(case, "zgbsv")
    print(name, kw, "info", got["info"].max(), "ipiv equal", np.array_equal(got["ipiv"], want["ipiv"]),
          "relmax", pc.relmax(got["x"], want["x"]), flush=True)
PY
echo "tiny rc=$?" >> gpurun_out/${TAG}_tiny.log
tail -8 gpurun_out/${TAG}_tiny.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/${TAG}_tests.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_v5.json 2> gpurun_out/${TAG}_bench_v5.err
echo "bench v5 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_v5.json'));print(d['ms_per_step'], d['kernels'])"
SZB_LIB=suzerain_b200/variants/libprof.so timeout -s KILL 200 python tools/prof_sync.py channel_192x96x192 18336 2>&1 | tee gpurun_out/${TAG}_prof.log
