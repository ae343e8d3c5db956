#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:invert_sync -s 1 -c 1 -o gpurun_out/r02_invert_sync_full \
    python tools/prof_invert.py channel_192x96x192 18336 > gpurun_out/r02_ncu_invert.log 2>&1
echo "invert capture rc=$?"; tail -2 gpurun_out/r02_ncu_invert.log
ls -la gpurun_out/r02_invert_sync_full.ncu-rep
