#!/bin/bash
# final single-GPU evidence: full GPU test suite, smoke, both bench arms as the driver runs them, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
echo "ref rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_reference.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['cpu_baseline']['cores'])"
timeout -s KILL 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench.err; python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','scaling','n_gpus')})
print('kernels', {k:round(v['ms'],3) for k,v in d['kernels'].items()}); print('solvers', {k:round(v['ms_per_step'],3) for k,v in d['solvers'].items()}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('fp64', d['roofline']['fp64']); print('cpu', d['cpu_baseline']['value']); print('clocks', d.get('clocks'))
PY
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_ncu.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
echo "launch list rc=$?"
