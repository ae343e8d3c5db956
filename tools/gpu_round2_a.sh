#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout -s KILL 1200 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 600 > gpurun_out/${TAG}_tests2.log 2>&1
echo "tests2 rc=$?"; tail -40 gpurun_out/${TAG}_tests2.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','scaling')})
print('kernels', d['kernels']); print('solvers', d['solvers']); print('e2e', d['e2e']); print('roofline fp64', d['roofline']['fp64'])
PY
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_ref.json | cut -c1-600
