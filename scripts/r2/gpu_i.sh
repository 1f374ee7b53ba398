#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("  ", d["case"][:70].ljust(70), d["payload_GBps"], d["ms_per_step"], d.get("hbm_frac_of_measured_copy_peak"))
    elif "rror" in l: print(l.rstrip()[:300])'
for env in "DDS_GATHER_GEOM_VAR=0" "DDS_GATHER_GEOM_VAR=8" "DDS_GATHER_GEOM_VAR=9" "DDS_GATHER_GEOM_VAR=2" "DDS_GATHER_GEOM_VAR=8 DDS_VAR_MINSEG=8" "DDS_GATHER_GEOM_VAR=6 DDS_GATHER_CTAS_PER_SM=2"; do
echo "== $env" | tee -a gpurun_out/r2i_configs.txt
env $env timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | python -c "$fmt" | grep -E "OVERLAP|^==" | tee -a gpurun_out/r2i_configs.txt
done
echo "== probe geom 8" | tee -a gpurun_out/r2i_configs.txt
DDS_GATHER_GEOM_VAR=8 timeout 300 python scripts/probes/queue_probe.py 2>&1 | grep -v Warn | head -18 | tee -a gpurun_out/r2i_configs.txt
