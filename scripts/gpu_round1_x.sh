#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg3_r1b.csv python scripts/bench_configs.py --cases cfg3 --scale 0.2 --steps 3 --warmup 2 > gpurun_out/ncu_cfg3b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dds_gather -s 20 -c 1 -o gpurun_out/prof_gather_var_r1b -f python scripts/bench_configs.py --cases cfg3 --scale 0.2 --steps 3 --warmup 2 > gpurun_out/ncu_cfg3_fullb.log 2>&1
tail -2 gpurun_out/ncu_cfg3_fullb.log
