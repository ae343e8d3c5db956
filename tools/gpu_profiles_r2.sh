#!/bin/bash
# Round-2 evidence: launch list of the bench command, full ncu captures of the dominant kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_ncu.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:invert_sync -s 1 -c 1 -o gpurun_out/r02_invert_sync_full \
    python tools/prof_invert.py channel_192x96x192 18336 > gpurun_out/r02_ncu_invert.log 2>&1
echo "invert capture rc=$?"; tail -2 gpurun_out/r02_ncu_invert.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:accumulate_kernel -s 1 -c 1 -o gpurun_out/r02_accumulate_full \
    python tools/prof_invert.py channel_192x96x192 18336 > gpurun_out/r02_ncu_acc.log 2>&1
echo "accumulate capture rc=$?"
ls -la gpurun_out/r02_*
