#!/bin/bash
# round-end evidence: launch list of the bench command (our kernels only) and DRAM traffic of the dominant kernel
mkdir -p gpurun_out
# our own kernels (the two cub calls of the library -- scan and radix sort -- share their names with torch's and are left out)
K='regex:^(correct_kernel|coverage_kernel|gather_kernel|kmer_count_kernel|len_to_u64_kernel|cost_key_kernel|table_.*|ctx_.*|model_tabs_kernel)$'
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches_r01_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_launch.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct \
    --clock-control none -k regex:correct_kernel -c 1 --csv --log-file gpurun_out/traffic_r01_final.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/final_ncu_traffic.log 2>&1
tail -8 gpurun_out/traffic_r01_final.csv
grep -c . gpurun_out/launches_r01_final.csv
