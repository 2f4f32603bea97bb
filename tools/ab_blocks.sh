#!/bin/bash
# residency sweep of the correction kernel with the default library (TALC_BLOCKS_PER_SM x 4 warps per SM)
mkdir -p gpurun_out
N=${1:-20000}
: > gpurun_out/ab_blocks.log
for b in 1 2 3 4; do
  echo "== blocks/SM $b" >> gpurun_out/ab_blocks.log
  TALC_BLOCKS_PER_SM=$b python tools/profile_case.py $N 2 >> gpurun_out/ab_blocks.log 2>&1
done
cat gpurun_out/ab_blocks.log
