#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dataset.py -m gpu -x -q 2>&1 | tail -4
for env in "DDS_L2_PERSIST=1" "DDS_L2_PERSIST=0"; do
echo "== $env"
env $env timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --configs cfg3,cfg4 2> gpurun_out/r2q_err.txt | python -c '
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "roof", d["roofline"]["frac"])
        for c in d["configs"]:
            print("  ", c["name"].ljust(28), "%8.1f" % c.get("value",0), "ms %.4f" % c.get("ms_per_step",0), "frac", round(c.get("roofline",{}).get("frac",0),3), "ser", round(c.get("serialized_ms_per_step",0),4), "ver", c.get("verified_rows"))'
grep -v CUDAEvent gpurun_out/r2q_err.txt | tail -3
done
