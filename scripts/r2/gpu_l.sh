#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2l_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err
tail -4 gpurun_out/r2l_bench_n1.err
python - <<'PY'
import json
for l in open("gpurun_out/r2l_bench_n1.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], d["ms_per_step_p10_p50_p90"], "ser", d["serialized_ms_per_step"], "roof", d["roofline"]["frac"], "ver", d["verified_rows"], d["mismatches"])
        print("e2e", {k:v for k,v in d["e2e"].items() if k!="path"})
        print("cpu", {k:v for k,v in d["cpu_baseline"].items() if k!="sample"})
        for c in d["configs"]:
            print("  ", c["name"].ljust(28), "%8.1f" % c.get("value",0), "ms %.4f" % c.get("ms_per_step",0), "frac", round(c.get("roofline",{}).get("frac",0),3), "ser", c.get("serialized_ms_per_step"), "ver", c.get("verified_rows"), "ref", (c.get("reference") or {}).get("value"))
            if c["name"] in ("per_sample_loop","ingest","prefetch_overlap"): print("      ", {k:v for k,v in c.items() if k not in ("workload","reference")})
PY
