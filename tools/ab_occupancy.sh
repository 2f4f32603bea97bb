#!/bin/bash
# A/B of the correction kernel's residency: same workload, libraries built with different TALC_MIN_BLOCKS
mkdir -p gpurun_out
N=${1:-20000}
echo "== default" > gpurun_out/ab.log
python tools/profile_case.py $N 2 >> gpurun_out/ab.log 2>&1
for mb in 6 8 12; do
  echo "== mb$mb" >> gpurun_out/ab.log
  TALC_LIB=talc_b200/_build/variants/lib_mb$mb.so TALC_BLOCKS_PER_SM=$mb python tools/profile_case.py $N 2 >> gpurun_out/ab.log 2>&1
done
cat gpurun_out/ab.log
