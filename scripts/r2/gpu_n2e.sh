#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --push > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -2 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_bench_n2.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("N 2 value", round(d["value"],1), d["fetch_mode"], "ms", round(d["ms_per_step"],4), "roof", round(d["roofline"]["frac"],3), "ver", d["verified_rows"], d["mismatches"])
        print("  push", d["push"] and (round(d["push"]["value"],1), round(d["push"]["ms_per_step"],4)), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["pageable_dst_value"],1))
PY
