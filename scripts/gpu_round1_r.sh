#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/bench_configs.py --cases demo,cfg5 --steps 20 --warmup 3 2>&1 | grep -E "^\{|rror" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['case'][:62].ljust(62), d['payload_GBps'], d['ms_per_step'])
    else: print(l.rstrip()[:200])" | head -6
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],3))"
