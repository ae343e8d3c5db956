#!/bin/bash
# 1-GPU run of both bench arms as the driver launches them (+ the e2e breakdown)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2u}
timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','scaling','n_gpus')})
print('kernels', {k:round(v['ms'],3) for k,v in d['kernels'].items()}); print('solvers', {k:round(v['ms_per_step'],3) for k,v in d['solvers'].items()}); print('e2e', d['e2e']); print('roofline', d['roofline']); print('cpu', d['cpu_baseline'])
PY
timeout -s KILL 300 python tools/diag_e2e.py 2>&1 | tail -12
