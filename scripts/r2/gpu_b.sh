#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for env in "DDS_SMEM_PLAN=1" "DDS_SMEM_PLAN=0" "DDS_SMEM_PLAN=1 DDS_GATHER_GEOM_S=3"; do
  echo "######## $env" | tee -a gpurun_out/r2b_timing.txt
  env $env timeout 300 python scripts/probes/timing_probe.py 2>&1 | grep -v Warning | tee -a gpurun_out/r2b_timing.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2b_pytest.txt
