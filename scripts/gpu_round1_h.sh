#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for g in 0 4 2 1; do
  echo "== GEOM(all) $g"
  DDS_GATHER_GEOM=$g timeout 600 python scripts/bench_configs.py --cases demo,cfg3,cfg4 --steps 20 --warmup 3 2>&1 | grep -E "^\{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['case'][:58].ljust(58), d['payload_GBps'], d['ms_per_step'])"
done
DDS_GATHER_GEOM=0 timeout 600 python scripts/bench_configs.py --cases cfg5 --steps 20 --warmup 3 2>&1 | grep -E "^\{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['case'][:58].ljust(58), d['payload_GBps'], d['ms_per_step'])" | head -3
