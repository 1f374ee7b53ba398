#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 scripts/probes/push_timing.py 2>&1 | grep -v "OMP\|\*\*\*\*" | tee gpurun_out/r2n2d_push_timing.txt
