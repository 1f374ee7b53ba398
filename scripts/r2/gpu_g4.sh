#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for bh in 1 4; do timeout 120 scripts/probes/gather4_probe $bh 2>&1 | tail -4 | tee -a gpurun_out/r2_gather4_probe.txt; done
