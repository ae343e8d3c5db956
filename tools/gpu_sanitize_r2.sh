#!/bin/bash
# Round-2 evidence: compute-sanitizer over the round-2 kernels, ncu captures of the new ones.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
# memcheck: the v5 fused invert (small and heavy-pivoting cases, both solvers), refinement, state exchange,
# bsplineop family, collect_references, drop-in wrappers
timeout -s KILL 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest -m gpu -q -x --timeout 1200 \
    tests/test_gpu_round2.py -k "zgbtrs or zaPxpby or exchange or set_refs or collect or bsplineop or dropin or sharded" \
    "tests/test_gpu_parity.py::test_invert_matches_oracle" > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r02_memcheck.log
# racecheck: shared-memory hazards of the fused invert (named barriers, mbarrier ring) and of collect_references
timeout -s KILL 1500 $CS --tool racecheck --error-exitcode 9 python -m pytest -m gpu -q -x --timeout 1200 \
    "tests/test_gpu_parity.py::test_invert_matches_oracle" "tests/test_gpu_round2.py::test_collect_references_matches_oracle" \
    "tests/test_gpu_round2.py::test_bsplineop_real_and_in_place_match_reference" -k "tiny or shape0 or 24-5" > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/r02_racecheck.log
# ncu: collect_references and the real-pencil bop kernel
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:collect_references_kernel -s 3 -c 1 -o gpurun_out/r02_collect_references_full \
    python tools/bench_aux.py > gpurun_out/r02_ncu_collect.log 2>&1
echo "collect capture rc=$?"
ls -la gpurun_out/r02_*
