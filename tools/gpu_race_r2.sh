#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
bash tools/gpu_try_variants.sh prev
timeout -s KILL 1500 $CS --tool racecheck --error-exitcode 9 python -m pytest -m gpu -q -x --timeout 1200 \
    "tests/test_gpu_parity.py::test_invert_matches_oracle" "tests/test_gpu_round2.py::test_collect_references_matches_oracle" \
    "tests/test_gpu_round2.py::test_bsplineop_real_and_in_place_match_reference" -k "tiny or shape0 or 24-5" > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -12 gpurun_out/r02_racecheck.log | cut -c1-220
