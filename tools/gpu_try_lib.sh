#!/bin/bash
# timing of the default library and of one variant, then the parity tests THROUGH the variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$1; TAG=${2:-r2v}
bash tools/gpu_try_variants.sh $V
SZB_LIB=suzerain_b200/variants/lib$V.so timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --timeout 300 -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests($V) rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
