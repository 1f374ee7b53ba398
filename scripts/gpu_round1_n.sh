#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/probes/sanitize_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "sanitize-ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_variants or sample_index" 2>&1 | tail -3
