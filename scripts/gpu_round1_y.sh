#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dds_gather -s 3 -c 1 -o gpurun_out/prof_gather_r1c -f python bench.py --samples 2000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_c.log 2>&1
tail -1 gpurun_out/ncu_full_c.log
