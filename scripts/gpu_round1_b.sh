#!/bin/bash
# bench with the stream fix, geometry sweep, ncu launch list + full capture of the gather kernel
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err; cat gpurun_out/bench_b.json
for g in 0 1 2 3 4 5; do for c in 1 2; do
  echo "GEOM $g CTAS_PER_SM $c"
  DDS_GATHER_GEOM=$g DDS_GATHER_CTAS_PER_SM=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  value', round(d['value'],1), 'per_launch_ms', round(d['roofline']['per_launch_ms'],4), 'frac', round(d['roofline']['frac'],3), d['config']['gather_geometry'])
"
done; done 2>&1 | tee gpurun_out/geom_sweep.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dds_gather -s 3 -c 2 -o gpurun_out/prof_gather_r1 -f python bench.py --samples 2000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
