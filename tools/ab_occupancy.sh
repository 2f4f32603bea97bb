#!/bin/bash
# A/B of the correction kernel's residency: libraries built with different TALC_MIN_BLOCKS (register caps);
# build them first: nvcc ... -DTALC_MIN_BLOCKS=3 -o talc_b200/_build/variants/lib_mb3.so (same for 4)
mkdir -p gpurun_out
: > gpurun_out/ab.log
for cfg in "default 2" "default 1" "mb3 3" "mb4 4"; do
  set -- $cfg
  echo "== lib $1 blocks/SM $2" >> gpurun_out/ab.log
  if [ "$1" = default ]; then L=""; else L="talc_b200/_build/variants/lib_$1.so"; fi
  TALC_LIB=$L TALC_BLOCKS_PER_SM=$2 TALC_DEBUG_CYCLES=1 python tools/profile_case.py 40000 2 2>&1 | grep -v "slow read" | tail -2 >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
