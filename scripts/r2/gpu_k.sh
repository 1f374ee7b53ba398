#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bindings.py tests/test_dataset.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2k_pytest.txt
for env in "DDS_DOORBELL=1" "DDS_DOORBELL=0"; do
echo "== $env" | tee -a gpurun_out/r2k_latency.txt
env $env timeout 300 python scripts/probes/latency_probe.py 2>&1 | tail -12 | head -4 | tee -a gpurun_out/r2k_latency.txt
done
