#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2f}
./tools/lat_bench > gpurun_out/${TAG}_lat.log 2>&1; cat gpurun_out/${TAG}_lat.log
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_v5.json 2> gpurun_out/${TAG}_bench_v5.err
echo "bench v5 rc=$?"; python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_v5.json'));print(d['ms_per_step'], d['kernels'])"
SZB_LIB=suzerain_b200/variants/libprof.so timeout -s KILL 200 python tools/prof_sync.py channel_192x96x192 18336 2>&1 | tee gpurun_out/${TAG}_prof.log
