#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r2w}; shift
bash tools/gpu_try_variants.sh "$@"
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --timeout 300 -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
SZB_LIB=suzerain_b200/variants/libprof.so timeout -s KILL 200 python tools/prof_sync.py channel_192x96x192 18336 2>&1 | tail -4 | tee gpurun_out/${TAG}_prof.log
