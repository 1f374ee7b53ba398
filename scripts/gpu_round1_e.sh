#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for pdl in 1 0; do
DDS_PDL=$pdl timeout 600 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('PDL=$pdl value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'per_launch_ms', round(d['roofline']['per_launch_ms'],4), 'frac', round(d['roofline']['frac'],3))
    else: print(l.rstrip())
"
done
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -3 gpurun_out/bench_e.err; cat gpurun_out/bench_e.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_e.json 2>&1; cat gpurun_out/bench_ref_e.json
