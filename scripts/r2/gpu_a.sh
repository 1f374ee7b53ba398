#!/bin/bash
# round 2, call A: parity tests on the new kernels + A/B of the variable-length path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.txt
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("  ", d["case"][:70].ljust(70), d["payload_GBps"], d["ms_per_step"], d.get("hbm_frac_of_measured_copy_peak"))
    elif "rror" in l: print(l.rstrip()[:300])'
for env in "DDS_SMEM_PLAN=1" "DDS_SMEM_PLAN=0" "DDS_SMEM_PLAN=1 DDS_S_MINSEG=4" "DDS_SMEM_PLAN=1 DDS_GATHER_GEOM_S=3" "DDS_SMEM_PLAN=1 DDS_GATHER_GEOM_S=2" "DDS_SMEM_PLAN=1 DDS_GATHER_GEOM_S=4"; do
  echo "== $env" | tee -a gpurun_out/r2a_configs.txt
  env $env timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | python -c "$fmt" | tee -a gpurun_out/r2a_configs.txt
done
echo "== fixed paths" | tee -a gpurun_out/r2a_configs.txt
timeout 600 python scripts/bench_configs.py --cases demo,cfg5 --steps 20 --warmup 3 2>&1 | python -c "$fmt" | tee -a gpurun_out/r2a_configs.txt
timeout 300 python scripts/probes/latency_probe.py 2>&1 | tail -12 | tee gpurun_out/r2a_latency.txt
