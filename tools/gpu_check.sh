#!/bin/bash
# tests + timing with per-read cycle tail + optional ncu capture of the correction kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TALC_DEBUG_CYCLES=1 python tools/profile_case.py ${1:-20000} 2 2>&1 | tail -12 | tee gpurun_out/check.log
TALC_BLOCKS_PER_SM=1 python tools/profile_case.py ${1:-20000} 2 2>&1 | tail -1 | tee -a gpurun_out/check.log
if [ -n "$2" ]; then
  ncu --set full --import-source on --clock-control none -k regex:correct_kernel -c 1 -f -o gpurun_out/$2 python tools/profile_case.py ${3:-4000} 1 > gpurun_out/ncu_$2.log 2>&1
  tail -2 gpurun_out/ncu_$2.log
fi
