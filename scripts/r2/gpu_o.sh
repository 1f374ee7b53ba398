#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dataset.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2o_pytest.txt
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("  ", d["case"][:70].ljust(70), d["payload_GBps"], d["ms_per_step"], d.get("hbm_frac_of_measured_copy_peak"))
    elif "rror" in l: print(l.rstrip()[:300])'
timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | python -c "$fmt" | grep -E "OVERLAP" | tee gpurun_out/r2o_configs.txt
timeout 300 python scripts/probes/queue_probe.py 2>&1 | grep -v Warn | head -18 | tee gpurun_out/r2o_queue_probe.txt
