#!/bin/bash
# round 2: ncu evidence (one GPU). Launch list of the bench command + full captures of the dominant kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# 1. launch list of bench.py (headline + cfg3 + cfg4), per-launch durations (cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_n1.csv \
   python bench.py --steps 4 --warmup 3 --repeats 1 --no-cpu-baseline --no-e2e --configs cfg3,cfg4 > gpurun_out/r2_ncu_bench.log 2>&1
# 2. --set full: fixed gather (config 2), variable gather (config 3, B=16384 and 4096), plan kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dds_gather -s 6 -c 1 -o gpurun_out/r2_gather_fixed_full \
   python bench.py --steps 4 --warmup 3 --repeats 1 --no-cpu-baseline --no-e2e --no-configs > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dds_gather -s 40 -c 1 -o gpurun_out/r2_gather_var_full \
   python scripts/bench_configs.py --cases cfg3 --steps 6 --warmup 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dds_plan -s 12 -c 1 -o gpurun_out/r2_plan_full \
   python scripts/bench_configs.py --cases cfg3 --steps 6 --warmup 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
for f in r2_gather_fixed_full r2_gather_var_full r2_plan_full; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.csv 2>/dev/null
done
