#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for z in 0 1; do
  if [ $z = 1 ]; then export DDS_NO_ZEROCOPY=1; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('NO_ZEROCOPY=$z value',round(d['value'],1),'e2e',round(d['e2e']['value'],2), 'ms', round(d['e2e']['ms_per_step'],3))
    elif 'rror' in l: print(l.rstrip()[:300])"
done
