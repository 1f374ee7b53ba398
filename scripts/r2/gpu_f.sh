#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("  ", d["case"][:70].ljust(70), d["payload_GBps"], d["ms_per_step"], d.get("hbm_frac_of_measured_copy_peak"))
    elif "rror" in l: print(l.rstrip()[:300])'
for env in "DDS_X=1" "DDS_GATHER_GEOM_VAR=6 DDS_GATHER_CTAS_PER_SM=2" "DDS_GATHER_GEOM_VAR=7 DDS_GATHER_CTAS_PER_SM=2" "DDS_GATHER_GEOM_VAR=5 DDS_GATHER_CTAS_PER_SM=2"; do
echo "== $env" | tee -a gpurun_out/r2f_configs.txt
env $env timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | python -c "$fmt" | grep -E "OVERLAP|^==" | tee -a gpurun_out/r2f_configs.txt
done
for env in "DDS_X=1" "DDS_GATHER_GEOM=6 DDS_GATHER_CTAS_PER_SM=2" "DDS_GATHER_GEOM=7 DDS_GATHER_CTAS_PER_SM=2"; do
echo "== fixed: $env" | tee -a gpurun_out/r2f_configs.txt
env $env timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --configs cfg5,cfg1 2>/dev/null | python -c '
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("   cfg2", round(d["value"],1), d["ms_per_step"], round(d["roofline"]["frac"],3), "ser", d["serialized_ms_per_step"])
        for c in d["configs"]: print("  ", c["name"].ljust(26), round(c["value"],1), round(c["ms_per_step"],4), round(c["roofline"]["frac"],3))' | tee -a gpurun_out/r2f_configs.txt
done
