#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dataset.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2m_pytest.txt
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --repeats 3 --no-e2e --no-cpu-baseline --configs prefetch 2>/dev/null | python -c '
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l)
        for c in d["configs"]: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in c.items() if k not in ("workload",)})' | tee -a gpurun_out/r2m_prefetch.txt
done
bash scripts/r2/gpu_san.sh
