#!/bin/bash
# N-GPU run of both bench arms as the driver launches them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r2k}
nvidia-smi -L | head -8; free -g | head -2; nproc
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err; python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench_${N}gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','scaling','n_gpus')}); print(d['config']['workload'])
print('kernels', {k:round(v['ms'],3) for k,v in d['kernels'].items()}); print('solvers', {k:round(v['ms_per_step'],3) for k,v in d['solvers'].items()}); print('e2e', d['e2e'])
PY
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_ref_${N}gpu.json 2> gpurun_out/${TAG}_ref_${N}gpu.err
echo "ref rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_ref_${N}gpu.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['cpu_baseline']['sample'][:120])"
