#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool" | tee -a gpurun_out/r2_sanitizer.txt
  timeout 1200 compute-sanitizer --tool $tool python scripts/probes/sanitize_small.py 2>&1 | grep -E "sanitize-ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard" | head -20 | tee -a gpurun_out/r2_sanitizer.txt
done
