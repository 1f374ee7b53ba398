#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | grep -E "^\{|rror" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['case'][:62].ljust(62), d['payload_GBps'], d['ms_per_step'])
    else: print(l.rstrip()[:200])"
