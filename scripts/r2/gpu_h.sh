#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dataset.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2h_pytest.txt
timeout 300 python scripts/probes/queue_probe.py 2>&1 | grep -v Warn | tee gpurun_out/r2h_queue_probe.txt
fmt='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("  ", d["case"][:70].ljust(70), d["payload_GBps"], d["ms_per_step"], d.get("hbm_frac_of_measured_copy_peak"))
    elif "rror" in l: print(l.rstrip()[:300])'
for env in "DDS_X=1" "DDS_GATHER_GEOM_VAR=5 DDS_GATHER_CTAS_PER_SM=2"; do
echo "== $env" | tee -a gpurun_out/r2h_configs.txt
env $env timeout 600 python scripts/bench_configs.py --cases cfg3,cfg4 --steps 20 --warmup 3 2>&1 | python -c "$fmt" | tee -a gpurun_out/r2h_configs.txt
done
