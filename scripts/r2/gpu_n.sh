#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dataset.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2n_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --configs prefetch,cfg3,cfg4,cfg1 2> gpurun_out/r2n_err.txt | python -c '
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], d["ms_per_step_p10_p50_p90"], "ser", d["serialized_ms_per_step"], "roof", d["roofline"]["frac"])
        for c in d["configs"]:
            if c["name"]=="prefetch_overlap": print({k:(round(v,3) if isinstance(v,float) else v) for k,v in c.items() if "hidden" in k or "exposed_prefetch" in k})
            else: print("  ", c["name"].ljust(28), "%8.1f" % c.get("value",0), "ms %.4f" % c.get("ms_per_step",0), "frac", round(c.get("roofline",{}).get("frac",0),3))' | tee gpurun_out/r2n_bench.txt
grep -v CUDAEvent gpurun_out/r2n_err.txt | tail -3
