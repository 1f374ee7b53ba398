#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2 | tee -a gpurun_out/r2_soak.txt
done
for i in 1 2 3 4; do
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "overlap" 2>&1 | tail -1 | tee -a gpurun_out/r2_soak.txt
done
