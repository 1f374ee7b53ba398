#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for m in copy push_stg push_tma; do timeout 120 scripts/probes/p2p_probe $m 2>&1 | tee -a gpurun_out/r2n2_push_probe.txt; done
