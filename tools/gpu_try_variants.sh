#!/bin/bash
# times the default library and every suzerain_b200/variants/lib<name>.so given as arguments (no tests)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo default; for i in 1 2; do timeout -s KILL 200 python tools/prof_invert.py channel_192x96x192 18336 2>&1 | grep "invert" | tail -1; done
for v in "$@"; do
  echo $v; for i in 1 2; do SZB_LIB=suzerain_b200/variants/lib$v.so timeout -s KILL 200 python tools/prof_invert.py channel_192x96x192 18336 2>&1 | grep "invert" | tail -1; done
done
