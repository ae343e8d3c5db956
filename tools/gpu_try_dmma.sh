#!/bin/bash
# quick check of a kernel change: timing first (two runs), then the parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2t}
for i in 1 2; do timeout -s KILL 200 python tools/prof_invert.py channel_192x96x192 18336 2>&1 | grep "invert" | tail -1; done
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --timeout 300 -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
if [ -f suzerain_b200/variants/libprof.so ]; then
SZB_LIB=suzerain_b200/variants/libprof.so timeout -s KILL 200 python tools/prof_sync.py channel_192x96x192 18336 2>&1 | tail -4 | tee gpurun_out/${TAG}_prof.log
fi
