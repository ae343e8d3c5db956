#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2j}
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/${TAG}_tests.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','scaling')})
print('kernels', {k:round(v['ms'],3) for k,v in d['kernels'].items()}); print('solvers', {k:(round(v['ms_per_step'],3), round(v['cpu']['value'],1)) for k,v in d['solvers'].items()}); print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'])
PY
