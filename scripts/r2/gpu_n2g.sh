#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2n2g.json 2> gpurun_out/r2n2g.err
grep real gpurun_out/r2n2g.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2n2g_ref.json 2> gpurun_out/r2n2g_ref.err
grep real gpurun_out/r2n2g_ref.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2n2g.json") if l.startswith("{")][0]
r=[json.loads(l) for l in open("gpurun_out/r2n2g_ref.json") if l.startswith("{")]
print("ours", round(d["value"],1), d["fetch_mode"], "push", d["push"], "e2e", round(d["e2e"]["value"],1), "configs", len(d["configs"]), "ver", d["verified_rows"])
print("ref lines", len(r), r and (round(r[0]["value"],1), r[0]["cpu_baseline"]["cores"], r[0]["config"]==d["config"]))
PY
