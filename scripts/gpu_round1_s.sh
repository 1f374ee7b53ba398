#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],3),'pair',round(d['roofline']['per_launch_event_pair_ms'],4))
    elif 'rror' in l: print(l.rstrip()[:300])"
done
