#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | grep -E "^\{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['case'][:62].ljust(62), d['payload_GBps'], d['ms_per_step'])"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -3 gpurun_out/bench_j.err; cat gpurun_out/bench_j.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_b.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dds_gather -s 3 -c 1 -o gpurun_out/prof_gather_r1b -f python bench.py --samples 2000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1
tail -2 gpurun_out/ncu_full_b.log
