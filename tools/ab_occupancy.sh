#!/bin/bash
# A/B of the correction kernel's residency: libraries built with different TALC_MIN_BLOCKS (register caps)
mkdir -p gpurun_out
N=${1:-20000}
: > gpurun_out/ab.log
for cfg in "default 4" "default 1" "mb1 1" "mb2 2" "mb2 1"; do
  set -- $cfg
  echo "== lib $1 blocks/SM $2" >> gpurun_out/ab.log
  if [ "$1" = default ]; then L=""; else L="talc_b200/_build/variants/lib_$1.so"; fi
  TALC_LIB=$L TALC_BLOCKS_PER_SM=$2 TALC_DEBUG_CYCLES=1 python tools/profile_case.py 20000 2 2>&1 | grep -v "slow read" | tail -2 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
