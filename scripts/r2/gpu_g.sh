#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/probes/queue_probe.py 2>&1 | grep -v Warn | tee gpurun_out/r2g_queue_probe.txt
