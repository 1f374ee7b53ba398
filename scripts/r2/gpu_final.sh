#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -3 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/r2_bench_ref_n1.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2_bench_n1.json") if l.startswith("{")][0]
r=[json.loads(l) for l in open("gpurun_out/r2_bench_ref_n1.json") if l.startswith("{")][0]
print("ours value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pageable", round(d["e2e"]["pageable_dst_value"],1), "| reference", round(r["value"],1), r["cpu_baseline"]["cores"], "threads of", r["host_cpus"], "| same config:", d["config"]==r["config"])
rc={c["name"]:c for c in r["configs"]}
for c in d["configs"]:
    if "roofline" in c: print("  ", c["name"].ljust(28), "%8.1f" % c["value"], "frac", round(c["roofline"]["frac"],3), "| ref arm", round(rc[c["name"]]["value"],1) if c["name"] in rc else None)
PY
