#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for env in "DDS_X=1" "DDS_GATHER_GEOM=1" "DDS_GATHER_GEOM=2" "DDS_GATHER_GEOM=7" "DDS_GATHER_GEOM=5 DDS_GATHER_CTAS_PER_SM=2"; do
  echo "== $env" | tee -a gpurun_out/r2n2f_geom.txt
  env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs --no-e2e 2>/dev/null | python -c '
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("   value", round(d["value"],1), "ms", round(d["ms_per_step"],4), d["ms_per_step_p10_p50_p90"], "ser", round(d["serialized_ms_per_step"],4), d["method"]["gather_geometry"])' | tee -a gpurun_out/r2n2f_geom.txt
done
