#!/bin/bash
# One GPU slot, everything that needs hardware (round 2): tests, A/B of the two execution shapes, bench, ncu.
# usage (on the box, from the repo root): bash tools/r2_gpu_session.sh [stage ...]   stages: tests ab bench rounds final dp ncu
# Every stage writes under gpurun_out/ and never stops the ones after it.
mkdir -p gpurun_out
STAGES="${@:-tests ab bench}"
for S in $STAGES; do
case $S in
tests)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=30 > gpurun_out/r2_tests.log 2>&1
  echo "== tests: $(tail -1 gpurun_out/r2_tests.log)"
  grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/r2_tests.log | head -20
  ;;
ab)
  # the two execution shapes on a mid-size workload (40 000 reads of config 2 at scale 0.02, table fits L2) and on
  # config 5; TALC_SPLIT / TALC_CTX / TALC_WALK_CAP select the shape
  for cfg in 2 5; do
    for shape in "0 16384 48" "1 16384 48" "1 16384 16" "1 32768 48" "1 8192 128"; do
      set -- $shape
      TALC_PROFILE_CONFIG=$cfg TALC_SPLIT=$1 TALC_CTX=$2 TALC_WALK_CAP=$3 timeout 600 python tools/profile_case.py 40000 2 \
        2>&1 | tail -1 | sed "s/^/cfg$cfg split=$1 ctx=$2 cap=$3: /"
    done
  done > gpurun_out/r2_ab.log 2>&1
  cat gpurun_out/r2_ab.log
  ;;
bench)
  TALC_SPLIT=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-table-load > gpurun_out/r2_bench_mono.json 2> gpurun_out/r2_bench_mono.err
  echo "== mono: $(python -c "import json;d=json.load(open('gpurun_out/r2_bench_mono.json'));print(d['value'],d['e2e']['value'],d.get('parity'))" 2>&1)"
  TALC_SPLIT=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.err
  echo "== split: $(python -c "import json;d=json.load(open('gpurun_out/r2_bench_split.json'));print(d['value'],d['e2e']['value'],d.get('parity'),d.get('table_load'),d['roofline'].get('peak_random'),d['roofline'].get('peak_random_dependent'))" 2>&1)"
  tail -2 gpurun_out/r2_bench_split.err
  ;;
dp)
  python tools/profile_dp.py 4096 300 > gpurun_out/r2_dp_plain.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:test_align_kernel -c 5 -o gpurun_out/r2_dp -f \
      python tools/profile_dp.py 4096 300 > gpurun_out/r2_ncu_dp.log 2>&1
  cat gpurun_out/r2_dp_plain.log
  ;;
final)
  # what the round-end driver runs, in the same order: tests, smoke, reference arm, own arm
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_tests_final.log 2>&1; tail -1 gpurun_out/r2_tests_final.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 300 gpurun_out/r2_bench_reference.json
  timeout 900 python bench.py --steps 6 --warmup 3 --cli-clock 1000000 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
  python -c "import json;d=json.load(open('gpurun_out/r2_bench_final.json'));print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity'],'table_load',d.get('table_load'),'cli',d.get('cli_clock'),'frac',d['roofline']['frac'],d['roofline'].get('frac_random'))" 2>&1
  tail -2 gpurun_out/r2_bench_final.err
  ;;
launches)
  # the launch list and the DRAM traffic of the bench command as it is now (two lanes; ncu serialises the launches)
  K='regex:^(correct_kernel|control_kernel|walk_kernel|coverage_kernel|gather_kernel|kmer_count_kernel|len_to_u64_kernel|cost_key_kernel|ctx_init_kernel|table_.*|ctx_.*|model_tabs_kernel|random_sector_kernel.*)$'
  timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-table-load > gpurun_out/r2_ncu_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-table-load > gpurun_out/r2_ncu_launch.log 2>&1
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,smsp__cycles_active.avg,sm__cycles_elapsed.max \
      --clock-control none -k regex:correct_kernel -c 1 --csv --log-file gpurun_out/r2_traffic_correct_kernel.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-table-load --lanes 1 > gpurun_out/r2_ncu_traffic.log 2>&1
  grep -c . gpurun_out/r2_launches.csv; tail -3 gpurun_out/r2_traffic_correct_kernel.csv | cut -c1-300
  ;;
ncu)
  K='regex:^(correct_kernel|control_kernel|walk_kernel|coverage_kernel|gather_kernel|kmer_count_kernel|len_to_u64_kernel|cost_key_kernel|ctx_init_kernel|table_.*|ctx_.*|model_tabs_kernel|random_sector_kernel.*)$'
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-table-load > gpurun_out/r2_ncu_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-table-load > gpurun_out/r2_ncu_launch.log 2>&1
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,smsp__cycles_active.avg,sm__cycles_elapsed.max \
      --clock-control none -k regex:correct_kernel -c 1 --csv --log-file gpurun_out/r2_traffic_correct_kernel.csv \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-table-load > gpurun_out/r2_ncu_traffic.log 2>&1
  grep -c . gpurun_out/r2_launches.csv
  TALC_PROFILE_CONFIG=2 ncu --set full --import-source on --clock-control none -k regex:correct_kernel -c 1 -o gpurun_out/r2_correct_kernel -f \
      python tools/profile_case.py 40000 1 > gpurun_out/r2_ncu_correct_kernel.log 2>&1
  for kern in control_kernel walk_kernel; do
    TALC_SPLIT=1 TALC_PROFILE_CONFIG=2 ncu --set full --import-source on --clock-control none -k regex:$kern -s 40 -c 1 -o gpurun_out/r2_$kern -f \
      python tools/profile_case.py 40000 1 > gpurun_out/r2_ncu_$kern.log 2>&1
  done
  ls -la gpurun_out/*.ncu-rep
  ;;
rounds)
  # per-round timing of the split shape (TALC_ROUND_DEBUG) on the two profile workloads, a few shapes
  for cfg in 2 5; do
    for shape in "16384 48 6 6 200000" "16384 48 0 6 200000" "32768 160 0 6 200000" "32768 160 0 6 1000000" "32768 160 0 6 0"; do
      set -- $shape
      echo "=== cfg$cfg ctx=$1 cap=$2 inlineInner=$3 inlineBorder=$4 pause=$5"
      TALC_ROUND_DEBUG=1 TALC_PROFILE_CONFIG=$cfg TALC_SPLIT=1 TALC_CTX=$1 TALC_WALK_CAP=$2 TALC_INLINE_INNER=$3 TALC_INLINE_BORDER=$4 TALC_PAUSE_CYCLES=$5 \
        timeout 600 python tools/profile_case.py 40000 1 2>&1 | grep -E "round|reads" | awk 'NR<=45 || /rounds\]/ || /reads/ || NR%10==0' | tail -80
    done
  done > gpurun_out/r2_rounds.log 2>&1
  grep -E "===|rounds\]|^reads" gpurun_out/r2_rounds.log
  ;;
esac
done
