#!/bin/bash
# Builds a variant of libsuzerain_b200.so with extra make variables / nvcc defines into
# suzerain_b200/variants/lib<name>.so (development only; select it with SZB_LIB=<path>).
#   tools/build_variant.sh <name> [make args...]      e.g.  tools/build_variant.sh prof PROF=1
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME=$1; shift
W=/tmp/szb_variant_$NAME
rm -rf $W; mkdir -p $W/suzerain_b200 $W/include
cp -r $ROOT/suzerain_b200/csrc $W/suzerain_b200/
cp $ROOT/include/*.h $W/include/
rm -f $W/suzerain_b200/csrc/*.o
make -C $W/suzerain_b200/csrc ../libsuzerain_b200.so "$@" > $W/build.log 2>&1 || { tail -30 $W/build.log; exit 1; }
mkdir -p $ROOT/suzerain_b200/variants
cp $W/suzerain_b200/libsuzerain_b200.so $ROOT/suzerain_b200/variants/lib$NAME.so
echo "built suzerain_b200/variants/lib$NAME.so"
