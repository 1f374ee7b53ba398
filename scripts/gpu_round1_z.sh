#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python scripts/bench_configs.py --cases cfg3 --steps 20 --warmup 3 2>&1 | grep -E "^\{|rror" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['case'][:66].ljust(66), d['payload_GBps'], d['ms_per_step'])
    else: print(l.rstrip()[:200])"
