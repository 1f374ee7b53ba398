#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -3 gpurun_out/bench_k.err; cat gpurun_out/bench_k.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step']); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])"
