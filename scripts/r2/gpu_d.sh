#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("  ", d["case"][:70].ljust(70), d["payload_GBps"], d["ms_per_step"], d.get("hbm_frac_of_measured_copy_peak"))
    elif "rror" in l: print(l.rstrip()[:300])'
for env in "DDS_PLAN_CARVEOUT=1" "DDS_PLAN_CARVEOUT=0" "DDS_SMEM_PLAN_MAX=8192"; do
echo "== $env" | tee -a gpurun_out/r2e_configs.txt
env $env timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | python -c "$fmt" | tee -a gpurun_out/r2e_configs.txt
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2e_pytest.txt
