#!/bin/bash
# On the GPU box: time QLT::run at ne120 x 1280 tracers for each variant library built by
# tools/ab_build.sh. Usage: tools/ab_run.sh NAME...
cd "$(dirname "$0")/.."
for n in "$@"; do
  CEDR_B200_LIB=$PWD/build/var/$n.so python tools/time_run.py qlt ne120x128x40 1280 > gpurun_out/ab_$n.log 2>&1
  echo "$n $(tail -n 1 gpurun_out/ab_$n.log)"
done
