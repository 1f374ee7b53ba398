#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("  ", d["case"][:70].ljust(70), d["payload_GBps"], d["ms_per_step"], d.get("hbm_frac_of_measured_copy_peak"))
    elif "rror" in l: print(l.rstrip()[:300])'
for env in "DDS_L2_PERSIST=1" "DDS_L2_PERSIST=0"; do
echo "== $env" | tee -a gpurun_out/r2p_configs.txt
env $env timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | python -c "$fmt" | grep -E "OVERLAP|rror" | tee -a gpurun_out/r2p_configs.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sample or multi_array" 2>&1 | tail -3
