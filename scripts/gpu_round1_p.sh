#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python scripts/probes/latency_probe.py 2>&1 | tail -10
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],3),'e2e',d['e2e']['value'], d['e2e']['ms_per_step'])"
