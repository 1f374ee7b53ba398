#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2j_pytest.txt
for env in "DDS_DOORBELL=1" "DDS_DOORBELL=0"; do
echo "== $env" | tee -a gpurun_out/r2j_latency.txt
env $env timeout 300 python scripts/probes/latency_probe.py 2>&1 | tail -12 | tee -a gpurun_out/r2j_latency.txt
done
timeout 600 python bench.py --steps 10 --warmup 3 --repeats 3 --no-e2e --configs persample,ingest,prefetch 2>/dev/null | python -c '
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l)
        for c in d["configs"]: print({k:v for k,v in c.items() if k not in ("workload",)})' | tee gpurun_out/r2j_extras.txt
