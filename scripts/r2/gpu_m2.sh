#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1; do
timeout 600 python bench.py --steps 10 --warmup 3 --repeats 3 --no-e2e --no-cpu-baseline --configs prefetch 2> gpurun_out/r2m_err.txt | python -c '
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l)
        for c in d["configs"]: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in c.items() if k not in ("workload",)})' | tee -a gpurun_out/r2m_prefetch.txt
tail -5 gpurun_out/r2m_err.txt
done
